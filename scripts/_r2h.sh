cd /root/repo
run() { tag=$1; cfg=$2; shift; shift
  env "$@" timeout 600 python scripts/bench_configs.py --config $cfg --steps 10 --out gpurun_out/r2h_$tag.json > gpurun_out/r2h_$tag.log 2>&1
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2h_$tag.json'))
    print('$tag', {k:d.get(k) for k in ('kernel','kernel_ms','device_ms','verify_ms','sort_and_rest_ms','candidates','n_hits','frac','frac_whole_pass','oracle_sample_equal')})
except Exception as e: print('$tag FAILED', e); print(open('gpurun_out/r2h_$tag.log').read()[-600:])
PY
}
run cfg4_base cfg4 X=1
run cfg4_f32 cfg4 MK_F32_MAX_FP=0.06
run cfg3_base cfg3 X=1
run cfg3_f32 cfg3 MK_F32_MAX_FP=0.06
run cfg4_f32_b16 cfg4 MK_F32_MAX_FP=0.1 MK_FILTER_BLOCKS=16384
MK_CUDA_LIB=$PWD/merkurio_b200/lib/libmerkurio_cuda_dbg.so timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4

cd /root/repo
echo "== debug-check build"; MK_CUDA_LIB=$PWD/merkurio_b200/lib/libmerkurio_cuda_dbg.so timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
echo "== TMA parity"; for m in 1 3; do MK_TMA=$m timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "random_single_length or bam4 or fixture or toy or double_buffered or record_boundaries" 2>&1 | tail -2; done
run() { tag=$1; cfg=$2; shift; shift
  env "$@" timeout 600 python scripts/bench_configs.py --config $cfg --steps 10 --out gpurun_out/r2i_$tag.json > gpurun_out/r2i_$tag.log 2>&1
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2i_$tag.json'))
    print('$tag', {k:d.get(k) for k in ('kernel','kernel_ms','device_ms','verify_ms','candidates','n_hits','frac','oracle_sample_equal')})
except Exception as e: print('$tag FAILED', e); print(open('gpurun_out/r2i_$tag.log').read()[-600:])
PY
}
run cfg4_base cfg4 X=1
run cfg4_tma3 cfg4 MK_TMA=3
run cfg3_base cfg3 X=1
run cfg3_tma1 cfg3 MK_TMA=1
run cfg3_tma2 cfg3 MK_TMA=2
run cfg3_tma3 cfg3 MK_TMA=3
for m in 0 1 2; do MK_TMA=$m python bench.py --steps 10 --no-e2e --no-configs --no-e2e-file --no-cpu-baseline --sustained-s 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 MK_TMA=$m', d['roofline']['kernel_ms'], d['roofline']['frac'], d['device_ms_per_step'])"; done
export MK_TMA=3
python scripts/bench_configs.py --config cfg4 --steps 2 > gpurun_out/r2i_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mk_scan_d16 -s 3 -c 1 -o gpurun_out/r2i_cfg4_tma -f python scripts/bench_configs.py --config cfg4 --steps 2 > gpurun_out/r2i_ncu.log 2>&1
tail -2 gpurun_out/r2i_ncu.log
unset MK_TMA
python scripts/bench_cli.py --config cfg2 --reads 12000000 --dir /dev/shm/mkcli --out gpurun_out/r2i_cli_cfg2.json 2>&1 | grep -E "^run|reader|engine setup" | head -12
rm -rf /dev/shm/mkcli

cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
run() { tag=$1; shift
  env "$@" timeout 600 python scripts/bench_configs.py --config cfg5 $EXTRA --steps 5 --out gpurun_out/r2d_$tag.json > gpurun_out/r2d_$tag.log 2>&1
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2d_$tag.json'))
    print('$tag', {k:d[k] for k in ('scan_ms','device_ms','verify_ms','candidates','n_hits','seeds','filter_bytes','sampled_queries_missing','hits_reverified','oracle_slice_hits')})
except Exception as e: print('$tag FAILED', e)
PY
}
EXTRA=--upper-queries run u_default X=1
EXTRA=--upper-queries run u_nogate MK_NO_GATE=1
EXTRA= run v_default X=1
EXTRA=--upper-queries run u_s1 MK_DUAL_SHAPE=1
EXTRA=--upper-queries run u_s2 MK_DUAL_SHAPE=2
EXTRA=--upper-queries run u_b24 MK_DUAL_BITS_PER_KEY=24
EXTRA=--upper-queries run u_b48 MK_DUAL_BITS_PER_KEY=48
python scripts/bench_configs.py --config cfg5 --upper-queries --steps 2 > gpurun_out/r2d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mk_scan_dual8 -s 3 -c 1 -o gpurun_out/r2d_dual8_v2 -f python scripts/bench_configs.py --config cfg5 --upper-queries --steps 2 > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log

cd /root/repo
run() { tag=$1; shift
  env "$@" timeout 600 python scripts/bench_configs.py --config cfg5 --upper-queries --steps 5 --out gpurun_out/r2c_$tag.json > gpurun_out/r2c_$tag.log 2>&1
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c_$tag.json'))
    print('$tag', {k:d[k] for k in ('scan_ms','device_ms','verify_ms','candidates','n_hits','seeds','filter_bytes','sampled_queries_missing','hits_reverified')})
except Exception as e: print('$tag FAILED', e)
PY
}
run t16o128 MK_DUAL_BITS_PER_KEY=16 MK_OFILTER_BITS=1048576
run t16o96 MK_DUAL_BITS_PER_KEY=16 MK_OFILTER_BITS=786432
run t16o64 MK_DUAL_BITS_PER_KEY=16 MK_OFILTER_BITS=524288
export MK_NO_TIER=1
python scripts/bench_configs.py --config cfg5 --upper-queries --steps 2 > gpurun_out/r2c_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mk_scan_dual8 -s 3 -c 1 -o gpurun_out/r2c_dual8_gate -f python scripts/bench_configs.py --config cfg5 --upper-queries --steps 2 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log

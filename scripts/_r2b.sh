cd /root/repo
run() { # tag, env..., -- args
  tag=$1; shift
  env "$@" timeout 600 python scripts/bench_configs.py --config cfg5 --upper-queries --steps 5 --out gpurun_out/r2b_$tag.json > gpurun_out/r2b_$tag.log 2>&1
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2b_$tag.json'))
    print('$tag', {k:d[k] for k in ('scan_ms','device_ms','verify_ms','candidates','n_hits','seeds','filter_bytes','sampled_queries_missing','hits_reverified')})
except Exception as e: print('$tag FAILED', e)
PY
}
run nt32 MK_NO_TIER=1
run nt16 MK_NO_TIER=1 MK_DUAL_BITS_PER_KEY=16
run nt8 MK_NO_TIER=1 MK_DUAL_BITS_PER_KEY=8
run t16 MK_DUAL_BITS_PER_KEY=16
run t12 MK_DUAL_BITS_PER_KEY=12
run t8 MK_DUAL_BITS_PER_KEY=8
run t16s1 MK_DUAL_BITS_PER_KEY=16 MK_DUAL_SHAPE=1
run t16s2 MK_DUAL_BITS_PER_KEY=16 MK_DUAL_SHAPE=2
run t16s3 MK_DUAL_BITS_PER_KEY=16 MK_DUAL_SHAPE=3
run nt16s1 MK_NO_TIER=1 MK_DUAL_BITS_PER_KEY=16 MK_DUAL_SHAPE=1

set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 > gpurun_out/r2a_parity.log
for v in "" "MK_NO_TIER=1" "MK_NO_TIER=1 MK_NO_GATE=1" "MK_NO_DUAL8=1 MK_NO_TIER=1"; do
  tag=$(echo "$v" | tr -d ' =' ); tag=${tag:-default}
  env $v timeout 600 python scripts/bench_configs.py --config cfg5 --out gpurun_out/r2a_cfg5_$tag.json > gpurun_out/r2a_cfg5_$tag.log 2>&1
  env $v timeout 600 python scripts/bench_configs.py --config cfg5 --upper-queries --out gpurun_out/r2a_cfg5u_$tag.json > gpurun_out/r2a_cfg5u_$tag.log 2>&1
done
tail -3 gpurun_out/r2a_parity.log
for f in gpurun_out/r2a_cfg5*.json; do echo $f; python -c "
import json,sys; d=json.load(open('$f')); print({k:d[k] for k in ('scan_ms','device_ms','verify_ms','candidates','n_hits','seeds','filter_bytes','table_build_s','frac_of_peak','hits_reverified','sampled_queries_missing','oracle_slice_hits')})"; done

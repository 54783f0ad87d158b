cd /root/repo
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv
python scripts/h2d_ceiling.py --out gpurun_out/r2g_h2d_ceiling.json 2>&1 | tail -12
python -c "
import json; d=json.load(open('gpurun_out/r2g_h2d_ceiling.json'))
print('\n'.join(d['nvidia-smi topo -m'][:14])); print('\n'.join(l for l in d['lscpu'] if any(k in l for k in ('Model name','Socket','NUMA','Core','CPU(s):')))); print(d['numactl -H'][:12])"
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --sustained-s 0 > gpurun_out/r2g_bench_n$n.json 2> gpurun_out/r2g_bench_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_n$n.json')); print($n, 'value', round(d['value']), 'e2e', {k: (round(v,3) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','h2d_ceiling_gb_per_s','frac_of_h2d_ceiling')}, 'clocks', d['clocks'])"
done
python -m pytest tests/test_gpu_cli.py -x -q -k "several_engines" 2>&1 | tail -3
python scripts/bench_cli.py --config cfg2 --reads 12000000 --gpus 8 --dir /dev/shm/mkcli --out gpurun_out/r2g_cli_cfg2_8gpu.json 2>&1 | tail -4
rm -rf /dev/shm/mkcli

cd /root/repo
(time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r2j_pytest.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err
(time python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err) 2> gpurun_out/r2j_bench.time
cat gpurun_out/r2j_pytest.log; grep -E "^\[bench\]" gpurun_out/r2j_bench_n1.err | cut -c1-400; cat gpurun_out/r2j_bench.time
S="--steps 3 --warmup 3 --no-e2e --no-configs --no-e2e-file --no-cpu-baseline --sustained-s 0"
python bench.py $S > gpurun_out/r2j_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches_bench.csv python bench.py $S > gpurun_out/r2j_ncu1.log 2>&1
python bench.py $S > gpurun_out/r2j_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"mk_scan_d16|mk_verify" -s 8 -c 2 -o gpurun_out/r2j_cfg2 -f python bench.py $S > gpurun_out/r2j_ncu2.log 2>&1
python scripts/bench_configs.py --config cfg5 --steps 2 > gpurun_out/r2j_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"mk_scan_dual8|mk_verify" -s 6 -c 2 -o gpurun_out/r2j_cfg5 -f python scripts/bench_configs.py --config cfg5 --steps 2 > gpurun_out/r2j_ncu3.log 2>&1
tail -2 gpurun_out/r2j_ncu2.log gpurun_out/r2j_ncu3.log

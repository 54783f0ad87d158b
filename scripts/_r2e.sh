cd /root/repo
scripts/micro/gather_bench 400 2>&1 | head -8 > gpurun_out/r2e_gather2.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2e_pytest.log 2>&1
(time python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err)  2> gpurun_out/r2e_bench.time
cat gpurun_out/r2e_gather2.txt; cat gpurun_out/r2e_pytest.log; tail -25 gpurun_out/r2e_bench.err; cat gpurun_out/r2e_bench.time

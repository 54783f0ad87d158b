cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -k "dual8_alphabet or dual_key or many_patterns" 2>&1 | tail -5
compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -x -q -k "toy or fixture or bam4 or dual_key or overflow or dual8_alphabet or window_scan or many_patterns or mixed_lengths" > gpurun_out/r2f_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -12 gpurun_out/r2f_memcheck.log

mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "golden" > gpurun_out/sanitizer_f.log 2>&1
grep -E "Invalid|misaligned|at 0x|by thread|Saved host|in /root|mrhbm::" gpurun_out/sanitizer_f.log | head -30

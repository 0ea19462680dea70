mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
timeout 400 compute-sanitizer --tool memcheck python __graft_entry__.py --smoke > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
tail -5 gpurun_out/pytest.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err

# round 2, call w (1 GPU): k_split_tma with launch-time facts folded (SPEC), lean tile loop; GPU suite + u64 / u64big / zipf benches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_w_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02_w_pytest.log | cut -c1-300
timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_w_u64.json 2> gpurun_out/r02_w_u64.err; echo "u64 rc=$?"
timeout 600 python bench.py --workload u64big --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_w_u64big.json 2> gpurun_out/r02_w_u64big.err; echo "u64big rc=$?"
timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r02_w_zipf.json 2> gpurun_out/r02_w_zipf.err; echo "zipf rc=$?"
python profiles/show.py gpurun_out/r02_w_u64.json gpurun_out/r02_w_u64big.json gpurun_out/r02_w_zipf.json | cut -c1-700
tail -n 3 gpurun_out/r02_w_*.err | cut -c1-300

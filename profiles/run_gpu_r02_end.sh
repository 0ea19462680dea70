# round 2, closing check (1 GPU): GPU suite, smoke, short benches of both workloads after the hook clean-up
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_end_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_end_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -n 2 | cut -c1-200
timeout 300 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_end_zipf.json 2> gpurun_out/r02_end_zipf.err; echo "zipf rc=$?"
timeout 300 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_end_u64.json 2> gpurun_out/r02_end_u64.err; echo "u64 rc=$?"
python profiles/show.py gpurun_out/r02_end_zipf.json gpurun_out/r02_end_u64.json | cut -c1-330

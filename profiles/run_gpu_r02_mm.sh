# round 2, call mm (1 GPU): the pipelined u64 sort also for bins whose key range is not known from their index (min / max pass): GPU suite + sequential-key bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_mm_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02_mm_pytest.log | cut -c1-300
timeout 600 python bench.py --workload u64seq --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_mm_u64seq.json 2> gpurun_out/r02_mm_u64seq.err; echo "u64seq rc=$?"
python profiles/show.py gpurun_out/r02_mm_u64seq.json | cut -c1-400

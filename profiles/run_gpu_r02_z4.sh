# round 2, call z4 (1 GPU): u64 sort with hoisted bucket shifts / no group scan for unique bins; level 1 after a read sweep (clean L2)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_z4_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02_z4_pytest.log | cut -c1-300
timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_z4_u64.json 2> gpurun_out/r02_z4_u64.err; echo "u64 rc=$?"
python profiles/show.py gpurun_out/r02_z4_u64.json | cut -c1-400
MRHBM_TUNE=$((64 + 128)) timeout 600 python bench.py --workload u64 --steps 4 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z4_u64_t.json 2> gpurun_out/r02_z4_u64_t.err; echo "rc=$?"
grep "split spans" gpurun_out/r02_z4_u64_t.err | tail -n 3

# round 2, call z6 (1 GPU): how far apart do level 1's CTAs start and end inside a shuffle? + the new parity test
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "neighbouring or sparse or duplicates or group_only" > gpurun_out/r02_z6_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02_z6_pytest.log | cut -c1-300
MRHBM_TUNE=64 timeout 600 python bench.py --workload u64 --steps 4 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z6_u64.json 2> gpurun_out/r02_z6_u64.err; echo "u64 rc=$?"
grep "level 1:" gpurun_out/r02_z6_u64.err | tail -n 4

# round 2, call z9 (1 GPU): u64 sort takes its bins from a ticket counter; GPU suite + u64 benches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_z9_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02_z9_pytest.log | cut -c1-300
MRHBM_TUNE=$((64 + 256)) timeout 600 python bench.py --workload u64 --steps 4 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z9_u64t.json 2> gpurun_out/r02_z9_u64t.err; echo "u64 rc=$?"
grep "u64 sort:" gpurun_out/r02_z9_u64t.err | tail -n 2 | cut -c60-200
timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_z9_u64.json 2> gpurun_out/r02_z9_u64.err; echo "u64 rc=$?"
timeout 600 python bench.py --workload u64big --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z9_u64big.json 2> gpurun_out/r02_z9_u64big.err; echo "u64big rc=$?"
python profiles/show.py gpurun_out/r02_z9_u64.json gpurun_out/r02_z9_u64big.json | cut -c1-420

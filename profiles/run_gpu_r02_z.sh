# round 2, call z (1 GPU): neighbour-window ranking in the u64 sort; is level 1 slower inside a shuffle because it is the first heavy kernel after an idle gap?
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_z_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02_z_pytest.log | cut -c1-300
timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_z_u64.json 2> gpurun_out/r02_z_u64.err; echo "u64 rc=$?"
MRHBM_TUNE=32 timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z_u64_t32.json 2> gpurun_out/r02_z_u64_t32.err; echo "u64 tune 32 rc=$?"
python profiles/show.py gpurun_out/r02_z_u64.json gpurun_out/r02_z_u64_t32.json | cut -c1-400

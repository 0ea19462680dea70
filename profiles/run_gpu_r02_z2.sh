# round 2, call z2 (1 GPU): %globaltimer spans of the two split levels inside a shuffle
mkdir -p gpurun_out
MRHBM_TUNE=64 timeout 600 python bench.py --workload u64 --steps 6 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z2_u64.json 2> gpurun_out/r02_z2_u64.err; echo "u64 rc=$?"
grep "split spans" gpurun_out/r02_z2_u64.err | tail -n 6
timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z2_u64b.json 2> gpurun_out/r02_z2_u64b.err; echo "u64 rc=$?"
python profiles/show.py gpurun_out/r02_z2_u64b.json | cut -c1-400

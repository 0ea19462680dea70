# round 2, call z8 (1 GPU): do the CTAs of the u64 sort and of the combiner end together?
mkdir -p gpurun_out
MRHBM_TUNE=$((64 + 256)) timeout 600 python bench.py --workload u64 --steps 4 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z8_u64t.json 2> gpurun_out/r02_z8_u64t.err; echo "u64 rc=$?"
grep "u64 sort:" gpurun_out/r02_z8_u64t.err | tail -n 3
MRHBM_TUNE=64 timeout 600 python bench.py --workload zipf32 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z8_zipft.json 2> gpurun_out/r02_z8_zipft.err; echo "zipf rc=$?"
grep "k_combine:" gpurun_out/r02_z8_zipft.err | tail -n 3
python profiles/show.py gpurun_out/r02_z8_zipft.json | cut -c1-300

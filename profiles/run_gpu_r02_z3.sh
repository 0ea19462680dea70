# round 2, call z3 (1 GPU): does staggering the co-resident CTAs of level 1 recover the 0.67 ms ncu sees?
mkdir -p gpurun_out
for st in 0 4 8 12 20; do
  MRHBM_TUNE=$((64 + st * 65536)) timeout 600 python bench.py --workload u64 --steps 6 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z3_u64_$st.json 2> gpurun_out/r02_z3_u64_$st.err; echo "stagger $st x 250 ns rc=$?"
  grep "split spans" gpurun_out/r02_z3_u64_$st.err | tail -n 2
done

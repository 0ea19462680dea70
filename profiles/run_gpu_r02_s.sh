# round 2, call s: prefetch-distance sweep of the combiner (MRHBM_TUNE bits 8-15), default back at 6 trips
mkdir -p gpurun_out
for pf in 0 3 4 5 8; do
MRHBM_TUNE=$((pf*256)) timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_s_zipf_pf$pf.json 2> gpurun_out/r02_s_zipf_pf$pf.err; echo "zipf pf=$pf rc=$?"
done
python profiles/show.py gpurun_out/r02_s_zipf_pf*.json 2>&1 | cut -c1-200

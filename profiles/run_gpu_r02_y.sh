# round 2, call y (1 GPU): where do the 0.12 ms between level 1 under ncu (0.67 ms) and inside a shuffle (0.80 ms) go?
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_y_launches_u64.csv python bench.py --workload u64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_y_l.log 2>&1; echo "launch list u64 rc=$?"
grep -E "k_split|k_sort|k_sample|k_exscan|k_region" gpurun_out/r02_y_launches_u64.csv | tail -n 14 | cut -d, -f5,12- | cut -c1-200
MRHBM_TUNE=2 timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_y_u64_exact.json 2> gpurun_out/r02_y_u64_exact.err; echo "u64 exact path rc=$?"
timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_y_u64.json 2> gpurun_out/r02_y_u64.err; echo "u64 rc=$?"
python profiles/show.py gpurun_out/r02_y_u64.json gpurun_out/r02_y_u64_exact.json | cut -c1-500

# round 2, call v (1 GPU): ncu captures for profiles/ -- launch lists of both workloads at full size, full sets of the hot kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v_launches_zipf32.csv python bench.py --workload zipf32 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_v_l1.log 2>&1; echo "launch list zipf rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v_launches_u64.csv python bench.py --workload u64 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_v_l2.log 2>&1; echo "launch list u64 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_split_tma|k_sort_reduce_u64" -s 9 -c 3 -o gpurun_out/r02_v_u64 python bench.py --workload u64 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_v_n1.log 2>&1; echo "ncu u64 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_combine|k_gtab_compact" -s 6 -c 2 -o gpurun_out/r02_v_zipf python bench.py --workload zipf32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_v_n2.log 2>&1; echo "ncu zipf rc=$?"
timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r02_v_u64.json 2> gpurun_out/r02_v_u64.err; echo "u64 rc=$?"
python profiles/show.py gpurun_out/r02_v_u64.json | cut -c1-600; tail -n 3 gpurun_out/r02_v_u64.err | cut -c1-300
ls -la gpurun_out/r02_v_*

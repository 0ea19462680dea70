# final evidence run of round 1: default bench (config 2), reference arm, config 3 bench, ncu launch
# list and full captures of the hot kernels
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; echo "rc=$?" >> gpurun_out/bench_r01_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_n1_reference.json 2>> gpurun_out/bench_r01_n1.err
timeout 900 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/bench_r01_n1_zipf32.json 2> gpurun_out/bench_r01_zipf.err; echo "rc=$?" >> gpurun_out/bench_r01_zipf.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01_zipf32.csv python bench.py --workload zipf32 --pairs 200000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_split|k_sort_reduce' -s 12 -c 3 -o gpurun_out/prof_r01_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_combine|k_agg_bins|k_scatter|k_hist' -s 8 -c 4 -o gpurun_out/prof_r01_final_zipf -f python bench.py --workload zipf32 --pairs 200000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_final_zipf.log 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
cut -c1-300 gpurun_out/bench_r01_n1.json; echo; cut -c1-300 gpurun_out/bench_r01_n1_reference.json; echo; cut -c1-300 gpurun_out/bench_r01_n1_zipf32.json; tail -n 2 gpurun_out/bench_r01_n1.err; tail -n 2 gpurun_out/bench_r01_zipf.err

mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "rc=$?" >> gpurun_out/bench_r01.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench_r01.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hist|k_scatter|k_sort_reduce' -s 9 -c 3 -o gpurun_out/prof_r01_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_final.log 2>&1
cat gpurun_out/bench_r01.json | cut -c1-400; tail -2 gpurun_out/bench_r01.err; cat gpurun_out/bench_r01_reference.json | cut -c1-300

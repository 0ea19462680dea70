mkdir -p gpurun_out
timeout 900 python bench.py --workload zipf32 --steps 5 --warmup 3 --cpu-sample 4000000 --e2e-steps 2 > gpurun_out/bench_r01_n1_zipf32.json 2> gpurun_out/bench_r01_zipf.err; echo "rc=$?" >> gpurun_out/bench_r01_zipf.err
cut -c1-600 gpurun_out/bench_r01_n1_zipf32.json; tail -n 2 gpurun_out/bench_r01_zipf.err | cut -c1-300

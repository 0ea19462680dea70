mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sort_reduce|k_split' -s 9 -c 3 -o gpurun_out/prof_r01_u -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_u.log 2>&1

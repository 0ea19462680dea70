mkdir -p gpurun_out
# one full-set capture per hot kernel (launch #2 of each, after warm-up), small step count
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hist|k_scatter|k_sort_reduce' -s 6 -c 3 -o gpurun_out/prof_r01_b -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_b.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_b.log
ls -la gpurun_out/

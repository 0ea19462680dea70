# round 2, call x (1 GPU): sort kernel ranks on 32-bit keys in a four-slot window; instruction counts of the u64 kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_x_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02_x_pytest.log | cut -c1-300
timeout 600 python bench.py --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_x_u64.json 2> gpurun_out/r02_x_u64.err; echo "u64 rc=$?"
python profiles/show.py gpurun_out/r02_x_u64.json | cut -c1-700
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_split_tma|k_sort_reduce_u64" -s 9 -c 3 -o gpurun_out/r02_x_u64 python bench.py --workload u64 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_x_n1.log 2>&1; echo "ncu u64 rc=$?"
tail -n 3 gpurun_out/r02_x_*.err | cut -c1-300

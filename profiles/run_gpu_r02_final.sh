# round 2, final capture (1 GPU): GPU suite, smoke, the driver's bench command and its reference arm, launch lists at full size, full ncu sets of the hot kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_final_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_final_smoke.log | cut -c1-300
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_reference_n1.json 2> gpurun_out/r02_final_reference_n1.err; echo "reference arm rc=$?"
python profiles/show.py gpurun_out/r02_final_bench_n1.json 2>&1 | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches_zipf32.csv python bench.py --workload zipf32 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_final_l1.log 2>&1; echo "launch list zipf rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches_u64.csv python bench.py --workload u64 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_final_l2.log 2>&1; echo "launch list u64 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_split_tma|k_sort_reduce_u64" -s 9 -c 3 -o gpurun_out/r02_final_u64 python bench.py --workload u64 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_final_n1.log 2>&1; echo "ncu u64 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_combine|k_gtab_compact" -s 6 -c 2 -o gpurun_out/r02_final_zipf python bench.py --workload zipf32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_final_n2.log 2>&1; echo "ncu zipf rc=$?"
ls -la gpurun_out/r02_final_* | cut -c30-

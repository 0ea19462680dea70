# round 2, call r: long keys batched, 3-word hash, prefetch 12 trips ahead; full default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "combiner or wordcount or zipf or nul or wider" > gpurun_out/r02_r_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_r_pytest.log | cut -c1-200
timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_r_zipf.json 2> gpurun_out/r02_r_zipf.err; echo "zipf rc=$?"
MRHBM_TUNE=4 timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_r_zipf_t4.json 2> gpurun_out/r02_r_zipf_t4.err; echo "zipf tune=4 rc=$?"
python profiles/show.py gpurun_out/r02_r_zipf.json gpurun_out/r02_r_zipf_t4.json 2>&1 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_combine -s 1 -c 1 -o gpurun_out/r02_r_combine python bench.py --workload zipf32 --pairs 200000000 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_r_ncu.log 2>&1; echo "ncu rc=$?"
#MRHBM_TUNE=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_combine -s 1 -c 1 -o gpurun_out/r02_r_combine_t4 python bench.py --workload zipf32 --pairs 200000000 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_r_ncu_t4.log 2>&1; echo "ncu t4 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_r_full.json 2> gpurun_out/r02_r_full.err; echo "full bench rc=$?"
python profiles/show.py gpurun_out/r02_r_full.json 2>&1 | cut -c1-400
tail -n 5 gpurun_out/r02_r_full.err | cut -c1-300

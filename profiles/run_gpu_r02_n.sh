# round 2, call n: pipelined global-table walk in the combiner + knock-out runs (MRHBM_TUNE 4 / 8 / 16: results invalid,
# timing only) that say where the combiner's time goes
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "combiner or wordcount or zipf or nul or wider" > gpurun_out/r02_n_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_n_pytest.log | cut -c1-200
timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_n_zipf.json 2> gpurun_out/r02_n_zipf.err; echo "zipf rc=$?"
for t in 4 8 16; do
MRHBM_TUNE=$t timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_n_zipf_t$t.json 2> gpurun_out/r02_n_zipf_t$t.err; echo "zipf tune=$t rc=$?"
done
timeout 300 python bench.py --workload u64 --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_n_u64.json 2> gpurun_out/r02_n_u64.err; echo "u64 rc=$?"
python profiles/show.py gpurun_out/r02_n_zipf.json gpurun_out/r02_n_zipf_t4.json gpurun_out/r02_n_zipf_t8.json gpurun_out/r02_n_zipf_t16.json gpurun_out/r02_n_u64.json 2>&1 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_combine -s 1 -c 1 -o gpurun_out/r02_n_combine python bench.py --workload zipf32 --pairs 200000000 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_n_ncu.log 2>&1; echo "ncu rc=$?"

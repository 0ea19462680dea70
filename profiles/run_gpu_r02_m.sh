# round 2, call m: global combiner table with 16-byte entries in buckets of four (one round trip per probe); contiguous tile ranges per split CTA
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "combiner or wordcount or zipf" > gpurun_out/r02_m_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_m_pytest.log | cut -c1-200
timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_m_zipf.json 2> gpurun_out/r02_m_zipf.err; echo "zipf rc=$?"
timeout 300 python bench.py --workload u64 --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_m_u64.json 2> gpurun_out/r02_m_u64.err; echo "u64 rc=$?"
python profiles/show.py gpurun_out/r02_m_zipf.json gpurun_out/r02_m_u64.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_combine -s 1 -c 1 -o gpurun_out/r02_m_combine python bench.py --workload zipf32 --pairs 200000000 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_m_ncu.log 2>&1; echo "ncu rc=$?"

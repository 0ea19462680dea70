# round 2, call c3 (1 GPU): combiner requests the first slot's key words together with the tags
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "combiner or zipf or wordcount or nul" > gpurun_out/r02_c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r02_c3_pytest.log | cut -c1-200
timeout 600 python bench.py --workload zipf32 --steps 8 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_c3_zipf.json 2> gpurun_out/r02_c3_zipf.err; echo "zipf rc=$?"
python profiles/show.py gpurun_out/r02_c3_zipf.json | cut -c1-300

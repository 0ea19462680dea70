# round 2, call u (1 GPU): pipelined tokeniser + bulk reduce in the job mirror; whole GPU suite, smoke, the driver's bench command
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_u_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02_u_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r02_u_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_u_smoke.log | cut -c1-300
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_u_full.json 2> gpurun_out/r02_u_full.err; echo "full bench rc=$?"
python profiles/show.py gpurun_out/r02_u_full.json 2>&1 | cut -c1-500
tail -n 3 gpurun_out/r02_u_full.err | cut -c1-300
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_u_ref.json 2> gpurun_out/r02_u_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02_u_ref.json

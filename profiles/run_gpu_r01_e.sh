mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>gpurun_out/bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print(round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in s.items()}, d['config'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sort_reduce' -s 3 -c 1 -o gpurun_out/prof_r01_e -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_e.log 2>&1

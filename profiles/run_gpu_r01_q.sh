mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
timeout 600 python bench.py --workload zipf32 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>gpurun_out/bench_zipf.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print('zipf 1e9', round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in s.items()}, d['config']['bins'], d['config']['parity_properties_ok'], d['config']['groups'])"
tail -2 gpurun_out/bench_zipf.err | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_combine' -s 1 -c 1 -o gpurun_out/prof_r01_combine -f python bench.py --workload zipf32 --pairs 200000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_q.log 2>&1

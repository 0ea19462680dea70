mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
timeout 600 python bench.py --workload zipf32 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>gpurun_out/bench_zipf.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print('zipf 1e9', round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in s.items()}, d['config']['bins'], d['config']['parity_properties_ok'], d['config']['groups'], d['roofline']['pipeline'])"
tail -n 2 gpurun_out/bench_zipf.err | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print('u64 1e8', round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in s.items()}, d['config']['parity_properties_ok'])"

mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -6 gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>gpurun_out/bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print(round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in s.items()}, d['config']['parity_properties_ok'], d['config']['bins'], d['clocks'])"
tail -2 gpurun_out/bench.err

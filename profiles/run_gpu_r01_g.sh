mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>gpurun_out/bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print(round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in s.items()}, d['config']['parity_properties_ok'], d['config']['bins'])"
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "golden or shared_prefix or zipf" > gpurun_out/sanitizer_g.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_g.log | tail -3

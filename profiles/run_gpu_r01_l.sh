mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
for n in 1000000000; do
timeout 600 python bench.py --workload zipf32 --pairs $n --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>gpurun_out/bench_zipf.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print('zipf n=$n', round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in s.items()}, d['config'])"
tail -2 gpurun_out/bench_zipf.err | cut -c1-300
done

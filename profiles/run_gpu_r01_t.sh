mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
for v in split; do
if [ $v = nosplit ]; then export MRHBM_NO_SPLIT=1; fi
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print('$v u64 1e8', round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in s.items()}, d['config']['parity_properties_ok'])"
done

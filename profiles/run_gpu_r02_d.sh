# round 2, call d: L2-resident global combiner table + unified optimistic path, one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_d_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r02_d_pytest.log | cut -c1-220
timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_d_u64.json 2> gpurun_out/r02_d_u64.err; echo "u64 rc=$?"
timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_d_zipf.json 2> gpurun_out/r02_d_zipf.err; echo "zipf rc=$?"
for f in u64 zipf; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_d_$f.json'))
    print('$f', round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3),'ms', 'launches', d['gpu_launches'], {k:round(v,3) for k,v in d['roofline']['stages_ms'].items()}, d['config'])
except Exception as e:
    print('$f failed', e); print(open('gpurun_out/r02_d_$f.err').read()[-1500:])
PY
done

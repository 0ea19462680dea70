# round 2, call g: combiner with hash tags in the shared table (straight-line probes); ncu of k_combine
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "combiner or wordcount or zipf or smoke" > gpurun_out/r02_g_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02_g_pytest.log | cut -c1-200
for t in 0; do
MRHBM_TUNE=$t timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_g_zipf$t.json 2> gpurun_out/r02_g_zipf$t.err; echo "zipf tune=$t rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_g_zipf$t.json'))
    print('zipf tune=$t', round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3),'ms', 'launches', d['gpu_launches'], {k:round(v,3) for k,v in d['roofline']['stages_ms'].items()}, d['config']['groups'], d['config']['pairs_after_combine'], d['config']['parity_properties_ok'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02_g_zipf$t.err').read()[-1500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_combine -s 1 -c 1 -o gpurun_out/r02_g_combine python bench.py --workload zipf32 --pairs 200000000 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_g_ncu.log 2>&1; echo "ncu rc=$?"

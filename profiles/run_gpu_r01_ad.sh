# multi-CTA exclusive scan: full GPU tests + default bench
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/pytest_ad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ad.log
tail -n 3 gpurun_out/pytest_ad.log
timeout -k 10 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ad_u64.json 2> gpurun_out/ad.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/ad_u64.json'))
print(round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3), d['gpu_launches'], {k:round(v,3) for k,v in d['roofline']['stages_ms'].items() if v>0.01})
PY
tail -n 3 gpurun_out/ad.err

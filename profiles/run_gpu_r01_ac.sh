# rolled FNV byte loop (instruction-cache fix) on config 3; full GPU tests
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/pytest_ac.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ac.log
tail -n 3 gpurun_out/pytest_ac.log
timeout -k 10 400 python bench.py --workload zipf32 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ac_zipf.json 2> gpurun_out/ac.err
timeout -k 10 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ac_u64.json 2>> gpurun_out/ac.err
for f in ac_zipf ac_u64; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/%s.json'%sys.argv[1]))
    print(sys.argv[1], round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stages_ms'].items() if v>0.01})
except Exception as ex: print(sys.argv[1], 'FAILED', ex)
PY
done
tail -n 5 gpurun_out/ac.err

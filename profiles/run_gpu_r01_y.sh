# bulk-copy (TMA) fed tile split: tests, sanitizer on the split tests, A/B benches on the same box
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/pytest_y.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_y.log
tail -n 8 gpurun_out/pytest_y.log
timeout -k 10 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "many_partitions or (uniform and 6000000)" --timeout 280 > gpurun_out/sanitizer_y.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer_y.log
tail -n 4 gpurun_out/sanitizer_y.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 0"
timeout -k 10 300 $B > gpurun_out/y_default.json 2> gpurun_out/y.err
MRHBM_NO_TMA_SPLIT=1 timeout -k 10 300 $B > gpurun_out/y_notma.json 2>> gpurun_out/y.err
timeout -k 10 300 $B > gpurun_out/y_default2.json 2>> gpurun_out/y.err
timeout -k 10 400 python bench.py --workload zipf32 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/y_zipf.json 2>> gpurun_out/y.err
for f in y_default y_notma y_default2 y_zipf; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/%s.json'%sys.argv[1]))
    print(sys.argv[1], round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stages_ms'].items() if v>0.01})
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
tail -n 3 gpurun_out/y.err

# pipelined u64 sort kernel + split variants: tests, then A/B benches on the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/pytest_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_x.log
tail -n 5 gpurun_out/pytest_x.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 0"
timeout 300 $B > gpurun_out/x_default.json 2> gpurun_out/x.err
MRHBM_NO_PIPELINED_SORT=1 timeout 300 $B > gpurun_out/x_oldsort.json 2>> gpurun_out/x.err
MRHBM_SPLIT_VARIANT=1 timeout 300 $B > gpurun_out/x_split1.json 2>> gpurun_out/x.err
timeout 300 $B > gpurun_out/x_default2.json 2>> gpurun_out/x.err
for f in x_default x_oldsort x_split1 x_default2; do python - "$f" <<'PY'
import json,sys
d=json.load(open('gpurun_out/%s.json'%sys.argv[1]))
print(sys.argv[1], round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stages_ms'].items() if v>0.01})
PY
done
tail -n 3 gpurun_out/x.err

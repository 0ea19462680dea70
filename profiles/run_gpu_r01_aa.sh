# micro-optimisations (lazy sub-bin hash, unrolled rank compare) + split tile/occupancy A/B
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/pytest_aa.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_aa.log
tail -n 4 gpurun_out/pytest_aa.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 0"
timeout -k 10 300 $B > gpurun_out/aa_cfg0.json 2> gpurun_out/aa.err
MRHBM_SPLIT_CFG=1 timeout -k 10 300 $B > gpurun_out/aa_cfg1.json 2>> gpurun_out/aa.err
MRHBM_SPLIT_CFG=2 timeout -k 10 300 $B > gpurun_out/aa_cfg2.json 2>> gpurun_out/aa.err
timeout -k 10 300 $B > gpurun_out/aa_cfg0b.json 2>> gpurun_out/aa.err
for f in aa_cfg0 aa_cfg1 aa_cfg2 aa_cfg0b; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/%s.json'%sys.argv[1]))
    print(sys.argv[1], round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stages_ms'].items() if v>0.01})
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
tail -n 3 gpurun_out/aa.err

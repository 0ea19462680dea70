# two-worker pipelined e2e vs serial e2e; large-tile 1-CTA/SM split configs
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -q -x --timeout 150 -k "uniform or many_partitions or golden" > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ab.log
tail -n 3 gpurun_out/pytest_ab.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline"
timeout -k 10 300 $B --e2e-steps 4 > gpurun_out/ab_e2e2.json 2> gpurun_out/ab.err
timeout -k 10 300 $B --e2e-steps 3 --e2e-serial > gpurun_out/ab_e2e1.json 2>> gpurun_out/ab.err
MRHBM_SPLIT_CFG=1 timeout -k 10 300 $B --e2e-steps 0 > gpurun_out/ab_cfg1.json 2>> gpurun_out/ab.err
MRHBM_SPLIT_CFG=2 timeout -k 10 300 $B --e2e-steps 0 > gpurun_out/ab_cfg2.json 2>> gpurun_out/ab.err
MRHBM_SPLIT_CFG=1 timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -q -x --timeout 150 -k "uniform or many_partitions" > gpurun_out/pytest_ab1.log 2>&1; echo "pytest cfg1 rc=$?" >> gpurun_out/pytest_ab1.log
tail -n 2 gpurun_out/pytest_ab1.log
for f in ab_e2e2 ab_e2e1 ab_cfg1 ab_cfg2; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/%s.json'%sys.argv[1]))
    e=d.get('e2e') or {}
    print(sys.argv[1], round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stages_ms'].items() if v>0.01}, 'e2e', round(e.get('value',0)/1e9,3), round(e.get('ms_per_step',0),2), e.get('groups_match'))
except Exception as ex: print(sys.argv[1], 'FAILED', ex)
PY
done
tail -n 5 gpurun_out/ab.err

mkdir -p gpurun_out
for v in 0 1; do
MRHBM_COMBINE_VARIANT=$v timeout 300 python bench.py --workload zipf32 --pairs 300000000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print('variant $v', {k:round(v,3) for k,v in s.items()}, d['config']['parity_properties_ok'])"
done

mkdir -p gpurun_out
for v in 0 1 2; do
  MRHBM_DEBUG_SCATTER=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print('variant=$v', {k:round(v,3) for k,v in s.items()})"
done

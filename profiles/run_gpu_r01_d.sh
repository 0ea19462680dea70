mkdir -p gpurun_out
for sh in 0 3 5; do
  MRHBM_CTR_SHIFT=$sh python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']; print('shift=$sh', round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in s.items()}, d['config']['parity_properties_ok'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sort_reduce' -s 3 -c 1 -o gpurun_out/prof_r01_d -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_d.log 2>&1

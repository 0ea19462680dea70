# round 2, call e: combiner with prefetch; ncu of k_combine (2e8-pair shape, one launch) for the stall picture
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_e_u64.json 2> gpurun_out/r02_e_u64.err; echo "u64 rc=$?"
timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_e_zipf.json 2> gpurun_out/r02_e_zipf.err; echo "zipf rc=$?"
for f in u64 zipf; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_e_$f.json'))
    print('$f', round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3),'ms', 'launches', d['gpu_launches'], {k:round(v,3) for k,v in d['roofline']['stages_ms'].items()})
except Exception as e:
    print('$f failed', e); print(open('gpurun_out/r02_e_$f.err').read()[-1500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_combine -s 1 -c 1 -o gpurun_out/r02_e_combine python bench.py --workload zipf32 --pairs 200000000 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_e_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -2

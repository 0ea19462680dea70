# round 2, call z5 (1 GPU): level 1 under ncu WITHOUT the cache flush between kernels (closer to the state inside a shuffle); zipf bench with the one-prefetch-per-CTA combiner
mkdir -p gpurun_out
timeout 900 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum --clock-control none -k regex:"k_split_tma|k_sort_reduce_u64" -c 12 --csv --log-file gpurun_out/r02_z5_nocacheflush.csv python bench.py --workload u64 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_z5_l.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_z5_nocacheflush.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(r[h.index('Kernel Name')][:44], r[h.index('Metric Name')], r[h.index('Metric Value')])
P
timeout 600 python bench.py --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_z5_zipf.json 2> gpurun_out/r02_z5_zipf.err; echo "zipf rc=$?"
python profiles/show.py gpurun_out/r02_z5_zipf.json | cut -c1-400

# round 2, call seq (1 GPU): the distribution the key-ordered fast path cannot take -- sequential u64 keys
mkdir -p gpurun_out
timeout 900 python bench.py --workload u64seq --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02_seq_u64seq.json 2> gpurun_out/r02_seq_u64seq.err; echo "u64seq rc=$?"
python profiles/show.py gpurun_out/r02_seq_u64seq.json | cut -c1-500
tail -n 3 gpurun_out/r02_seq_u64seq.err | cut -c1-300

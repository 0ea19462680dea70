# round 2, call t (2 GPUs): interleaved level-2 tile streams, asynchronous combiner; multi_gpu_check + benches
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/r02_t_multi$N.log 2>&1; echo "multi rc=$?"
grep -E "MULTI_GPU_CHECK|ok:|rc=|Error|error|assert" gpurun_out/r02_t_multi$N.log | head -20 | cut -c1-300
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload u64 --steps 20 --warmup 3 --e2e-steps 0 > gpurun_out/r02_t_u64_n$N.json 2> gpurun_out/r02_t_u64_n$N.err; echo "u64 rc=$?"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload u64big --steps 10 --warmup 3 --e2e-steps 0 > gpurun_out/r02_t_u64big_n$N.json 2> gpurun_out/r02_t_u64big_n$N.err; echo "u64big rc=$?"
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --workload zipf32 --steps 5 --warmup 3 --e2e-steps 0 > gpurun_out/r02_t_zipf_n$N.json 2> gpurun_out/r02_t_zipf_n$N.err; echo "zipf rc=$?"
python profiles/show.py gpurun_out/r02_t_u64_n$N.json gpurun_out/r02_t_u64big_n$N.json gpurun_out/r02_t_zipf_n$N.json
tail -n 3 gpurun_out/r02_t_*_n$N.err | cut -c1-300

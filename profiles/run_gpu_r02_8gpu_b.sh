# round 2, second 8-GPU call (final kernels: ticket counters, clump-aware sub-bins, long-key store): multi_gpu_check + the driver's bench command
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/r02_8b_multi$N.log 2>&1; echo "multi rc=$?"
grep -E "MULTI_GPU_CHECK|ok:|rc=|Error|error|assert" gpurun_out/r02_8b_multi$N.log | head -20 | cut -c1-300
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_8b_full_n$N.json 2> gpurun_out/r02_8b_full_n$N.err; echo "driver command rc=$?"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload u64 --steps 20 --warmup 5 --e2e-steps 0 --no-parity > gpurun_out/r02_8b_u64_n$N.json 2> gpurun_out/r02_8b_u64_n$N.err; echo "u64 weak rc=$?"
python profiles/show.py gpurun_out/r02_8b_full_n$N.json gpurun_out/r02_8b_u64_n$N.json | cut -c1-900
tail -n 3 gpurun_out/r02_8b_*_n$N.err | cut -c1-300

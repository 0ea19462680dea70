# round 2, 4-GPU call (budget-bound, short): multi_gpu_check + the headline block of the bench
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/r02_4_multi$N.log 2>&1; echo "multi rc=$?"
grep -E "MULTI_GPU_CHECK|rc=|Error|error|assert" gpurun_out/r02_4_multi$N.log | head -8 | cut -c1-300
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --only-headline --e2e-steps 1 > gpurun_out/r02_4_head_n$N.json 2> gpurun_out/r02_4_head_n$N.err; echo "headline rc=$?"
python profiles/show.py gpurun_out/r02_4_head_n$N.json | cut -c1-600

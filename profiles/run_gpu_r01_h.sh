mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/multi2.log 2>&1; echo "rc=$?" >> gpurun_out/multi2.log
grep -vE "^\[W|^W0|Warning|warn" gpurun_out/multi2.log | tail -25

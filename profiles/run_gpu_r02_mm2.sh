# round 2, call mm2 (2 GPUs): multi_gpu_check with the min / max variant of the u64 sort on the exact multi-GPU path
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/r02_mm2_multi$N.log 2>&1; echo "multi rc=$?"
grep -E "MULTI_GPU_CHECK|ok:|rc=|Error|error|assert" gpurun_out/r02_mm2_multi$N.log | head -14 | cut -c1-200

# round 2, call lk (2 GPUs): keys longer than a record slot (host side store) -- GPU suite incl. the 2-rank check
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_lk_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 30 gpurun_out/r02_lk_pytest.log | cut -c1-300
N=$(nvidia-smi -L | wc -l)
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/r02_lk_multi$N.log 2>&1; echo "multi rc=$?"
grep -E "MULTI_GPU_CHECK|ok:|rc=|Error|error|assert" gpurun_out/r02_lk_multi$N.log | head -20 | cut -c1-300

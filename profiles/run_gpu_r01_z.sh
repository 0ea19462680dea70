# 2-GPU validation of the exact bulk-copy split + multi-segment pipelined sort: correctness check, then bench
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/multi$N.log 2>&1; echo "rc=$?" >> gpurun_out/multi$N.log
grep -E "MULTI_GPU_CHECK|rc=|Error|error" gpurun_out/multi$N.log | head -8 | cut -c1-300
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
grep -E "Error|error|rc=" gpurun_out/bench_n$N.err | head -5 | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['n_gpus'], round(d['value']/1e9,2),'Gp/s', d['ms_per_step'], {k:round(v,3) for k,v in d['roofline']['stages_ms'].items()}, d['e2e'], d['config']['parity_properties_ok'])"

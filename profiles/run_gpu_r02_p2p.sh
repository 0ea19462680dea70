# FIRST THING for the next round (needs >= 2 GPUs): validate the experimental fused split -> peer-memory
# exchange (MRHBM_P2P=1) against the oracle, then A/B it against the NCCL exchange on the same box.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
MRHBM_P2P=1 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/multi_p2p$N.log 2>&1; echo "rc=$?" >> gpurun_out/multi_p2p$N.log
grep -E "MULTI_GPU_CHECK|rc=|Error|error" gpurun_out/multi_p2p$N.log | head -8 | cut -c1-300
for mode in 0 1; do
  MRHBM_P2P=$mode timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 0 > gpurun_out/bench_p2p${mode}_n$N.json 2> gpurun_out/bench_p2p${mode}_n$N.err; echo "rc=$?" >> gpurun_out/bench_p2p${mode}_n$N.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_p2p${mode}_n$N.json')); print('p2p=$mode', d['n_gpus'], round(d['value']/1e9,2),'Gp/s', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stages_ms'].items()}, d['config']['parity_properties_ok'])"
  tail -n 2 gpurun_out/bench_p2p${mode}_n$N.err | cut -c1-300
done

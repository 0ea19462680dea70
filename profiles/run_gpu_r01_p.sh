mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for ch in default 32; do
  if [ "$ch" != "default" ]; then export NCCL_MIN_P2P_NCHANNELS=$ch NCCL_MAX_P2P_NCHANNELS=$ch; fi
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('p2p channels=$ch', d['n_gpus'], round(d['value']/1e9,2),'Gp/s', {k:round(v,3) for k,v in d['roofline']['stages_ms'].items()})"
done

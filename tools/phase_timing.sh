# debug helper (GPU box, 2 GPUs): phase split of one step at N=2
UD_PHASE_TIMING=1 UD_DDP_DEBUG=1 UD_NCCL_MAX_CTAS=0 UD_DDP_SIDE_CTAS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | grep "phase\|^{\|timeline" | cut -c1-200

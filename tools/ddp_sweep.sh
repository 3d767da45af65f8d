python -m pytest tests/test_kernels_gpu.py -q -k "flat_helpers" 2>&1 | tail -2
for cfg in "0 0" "8 64" "4 32" "16 128" "2 16"; do
  set -- $cfg
  echo "== NCCL_MAX_CTAS=$1 SIDE_CTAS=$2"
  UD_NCCL_MAX_CTAS=$1 UD_DDP_SIDE_CTAS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['ms_per_step'], j['clocks'])"
done

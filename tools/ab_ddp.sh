for i in 1 2; do
for cfg in "base" "B200SR_NO_NCCL_REG=1"; do
  if [ "$cfg" = "base" ]; then envs=""; else envs="$cfg"; fi
  r=$(env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2953$i bench.py --gpus $NG --steps 30 --warmup 5 --no-variants 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['ddp']['nccl_user_buffer_registration'])")
  echo "N=$NG $cfg : $r"
done
done

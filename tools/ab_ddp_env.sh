# A/B of environment settings for the data-parallel bench inside one gpurun call. Usage: NG=8 bash tools/ab_ddp_env.sh "base" "NCCL_MAX_CTAS=4" ...
i=0
for cfg in "$@"; do
  i=$((i+1))
  if [ "$cfg" = "base" ]; then envs=""; else envs="$cfg"; fi
  r=$(env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus $NG --steps 30 --warmup 5 --no-variants 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])")
  echo "N=$NG $cfg : $r"
done

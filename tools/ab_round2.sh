for i in 1 2; do
for cfg in "base" "B200SR_UNPACK_DIRECT_MIN=100000" "B200SR_NO_FUSED_BN=1" "B200SR_UNPACK_DIRECT_MIN=100000 B200SR_NO_FUSED_BN=1"; do
  if [ "$cfg" = "base" ]; then envs=""; else envs="$cfg"; fi
  r=$(env $envs python bench.py --steps 30 --warmup 5 --no-variants --no-gpu-baseline --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])")
  echo "$cfg : $r"
done
done

# A/B inside one gpurun call: env configurations of bench.py (ms per step resident / e2e). Usage: bash tools/ab_round2.sh "cfgA" "cfgB" ...
for i in 1 2; do
for cfg in "$@"; do
  if [ "$cfg" = "base" ]; then envs=""; else envs="$cfg"; fi
  r=$(env $envs python bench.py --steps 30 --warmup 5 --no-variants --no-gpu-baseline --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])")
  echo "$cfg : $r"
done
done

#!/bin/bash
# A/B the train step under two environments, interleaved, in ONE gpurun call (box-to-box variance is ~5%).
# usage: tools/ab_bench.sh "ENV_A=1" "ENV_B=1" [reps]
A="$1"; B="$2"; N="${3:-3}"
for i in $(seq $N); do
  for v in "$A" "$B"; do
    r=$(env $v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])")
    echo "[$v] ms_per_step e2e: $r"
  done
done

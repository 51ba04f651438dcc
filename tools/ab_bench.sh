#!/bin/bash
# A/B the train step under several environments, interleaved, in ONE gpurun call (box-to-box variance is ~5%).
# usage: tools/ab_bench.sh reps "ENV_A=1" "ENV_B=1 ENV_C=2" ...
N="$1"; shift
for i in $(seq $N); do
  for v in "$@"; do
    r=$(env $v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']))")
    echo "[$v] ms_per_step e2e: $r"
  done
done

#!/bin/bash
# sweep the k_eval_obs tuning knobs (observations per thread, CTAs per SM) on the bench workload
for cfg in "1 4 1" "1 3 1" "2 4 2" "2 3 2" "4 3 4" "1 4 4" "2 3 4"; do
  set -- $cfg
  out=$(KBA_EVAL_PER_JAC=$1 KBA_EVAL_MIN_BLOCKS=$2 KBA_EVAL_PER_COST=$3 python bench.py --steps 2 --warmup 3 --batch 148 --cpu-sample 0 2>&1 | tail -1)
  echo "jac_per=$1 min_blocks=$2 cost_per=$3 :: $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f  jac_ms %.4f  frac %.3f' % (d['value'], d['roofline']['launch_ms_mean'], d['roofline']['frac']))")"
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log
tail -15 gpurun_out/r2_t7.log | cut -c1-400
for cm in 0 1; do
echo "copy mode $cm"
KBA_FUSED_COPY=$cm KBA_LIB_PATH=$PWD/limo_b200/libkba_b200_prof.so timeout 300 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 2>&1 | grep "kba prof" | head -2
KBA_FUSED_COPY=$cm timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.1f e2e %.1f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
timeout 600 python scripts/eval_sweep.py 2>&1 | tail -16

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t8.log
tail -8 gpurun_out/r2_t9.log | cut -c1-400
KBA_LIB_PATH=$PWD/limo_b200/libkba_b200_prof.so timeout 300 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 2>&1 | grep "kba prof" | head -2
for tj in 8; do
KBA_EVAL_TILES_JAC=$tj timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tiles $tj: value %.1f e2e %.1f ms/step %.1f jac_ms %.4f frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_mean'], d['roofline']['frac']))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches9.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu9.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches9.csv

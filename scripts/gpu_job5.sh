#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t5.log
tail -8 gpurun_out/r2_t5.log
for ov in 1 0; do
KBA_OVERLAP=$ov timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-sub > gpurun_out/r2_b5_ov$ov.json 2> gpurun_out/r2_b5_ov$ov.err; echo "bench overlap=$ov rc=$?"; tail -3 gpurun_out/r2_b5_ov$ov.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_b5_ov$ov.json").read().strip().splitlines()[-1])
    print("overlap=$ov value %.1f e2e %.1f ms/step %.1f jac_ms %.4f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_ms_mean"], d["roofline"]["frac"]))
except Exception as e:
    print("failed", e)
PY
done
KBA_LIB_PATH=$PWD/limo_b200/libkba_b200_prof.so timeout 300 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 2>&1 | grep "kba prof" | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches5.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu5.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches5.csv
# one pass of every kernel, full metric set (launches 60.. = first passes after the first trimming round)
KBA_OVERLAP=0 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 60 --launch-count 14 -f -o gpurun_out/prof_r02_a python bench.py --steps 1 --warmup 0 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncufull5.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_r02_a.ncu-rep > gpurun_out/r2_ncu_summary5.txt 2>&1; head -60 gpurun_out/r2_ncu_summary5.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log
tail -8 gpurun_out/r2_t4.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_b4.json 2> gpurun_out/r2_b4.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_b4.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_b4.json").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f ms/step %.1f jac_ms %.4f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_ms_mean"], d["roofline"]["frac"]))
    print(json.dumps(d["cpu_baseline"]))
    print(json.dumps(d["sub_records"], indent=1)[:6000])
except Exception as e:
    print("failed", e)
PY
KBA_LIB_PATH=$PWD/limo_b200/libkba_b200_prof.so timeout 300 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 2>&1 | grep "kba prof" | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches4.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu4.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches4.csv

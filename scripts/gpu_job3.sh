#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
tail -12 gpurun_out/r2_t3.log
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 1 > gpurun_out/r2_b3_fused.json 2> gpurun_out/r2_b3_fused.err
python - <<'PY'
import json
for n in ("fused",):
    try:
        d = json.loads(open("gpurun_out/r2_b3_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.1f e2e %.1f ms/step %.1f jac_ms %.4f frac %.3f conv %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_ms_mean"], d["roofline"]["frac"], d["config"]["all_windows_converged"]))
    except Exception as e:
        print(n, "failed", e)
PY
KBA_LIB_PATH=$PWD/limo_b200/libkba_b200_prof.so timeout 300 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --batch 148 --in-flight 1 2>&1 | grep "kba prof" | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches3.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --batch 148 --in-flight 1 > gpurun_out/r2_ncu3.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches3.csv

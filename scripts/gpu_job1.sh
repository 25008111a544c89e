#!/bin/bash
# round 2, first GPU pass over the fused small-window path: parity tests, A/B bench against the round-1 kernels, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
tail -15 gpurun_out/r2_t1.log
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 1 > gpurun_out/r2_b1_fused.json 2> gpurun_out/r2_b1_fused.err
KBA_FUSED=0 timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 1 > gpurun_out/r2_b1_old.json 2> gpurun_out/r2_b1_old.err
python - <<'PY'
import json
for n in ("fused", "old"):
    try:
        d = json.loads(open("gpurun_out/r2_b1_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.1f e2e %.1f ms/step %.1f jac_ms %.4f frac %.3f conv %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_ms_mean"], d["roofline"]["frac"], d["config"]["all_windows_converged"]))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches1.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --batch 148 --in-flight 1 > gpurun_out/r2_ncu1.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches1.csv

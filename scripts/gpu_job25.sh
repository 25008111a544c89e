#!/bin/bash
# validation of the final build (strided grids by default): GPU suite, smoke, the default bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu --durations=5 ) > gpurun_out/r2_t25.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_t25.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench25.json 2> gpurun_out/r2_bench25.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench25.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench25.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f (seq %.1f) ms/step %.1f lin frac %.3f jac-alone frac %.3f launches %d" % (d["value"], d["e2e"]["value"], d["e2e"]["sequential_value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline_jacobian_kernel"]["frac"], d["gpu_launches"]))
for b in d["sub_records"]["config2_batch"]: print(json.dumps(b)[:300])
print(json.dumps(d["sub_records"]["config3"])[:600])
print(json.dumps(d["sub_records"]["config4_lidar"])[:400])
print(json.dumps(d["sub_records"]["config5"])[:400])
print(json.dumps(d["cpu_baseline"])[:500])
PY

#!/bin/bash
# pipelined completion polling: GPU suite, default bench (batch-1 latency in sub_records), 2 solves under the launch list
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_t18.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_t18.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench18.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench18.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f (seq %.1f) ms/step %.1f lin frac %.3f jac-alone frac %.3f launches %d" % (d["value"], d["e2e"]["value"], d["e2e"]["sequential_value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline_jacobian_kernel"]["frac"], d["gpu_launches"]))
print(json.dumps(d["sub_records"]["config2_batch"])[:900])
print(json.dumps(d["sub_records"]["config5"])[:400])
PY

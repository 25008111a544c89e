#!/bin/bash
# final validation of the round: GPU suite (incl. the new tests), smoke, default bench, ncu --set full of one pass of the final build
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu --durations=5 ) > gpurun_out/r2_t19.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_t19.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench19.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench19.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f (seq %.1f) ms/step %.1f lin frac %.3f jac-alone frac %.3f launches %d" % (d["value"], d["e2e"]["value"], d["e2e"]["sequential_value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline_jacobian_kernel"]["frac"], d["gpu_launches"]))
print(json.dumps(d["sub_records"]["config2_batch"][0]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -s 130 -c 22 -o gpurun_out/r2_full19 -f python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu19_full.log 2>&1
ls -la gpurun_out/r2_full19.ncu-rep

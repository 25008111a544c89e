#!/bin/bash
# device-driven solve loop (CUDA graph with a conditional WHILE node), tile-packed A from k_sred_reduce, descriptor-only tiles in
# k_linearize, hoisted loads in k_backsub_v, 1024-thread k_solve_begin: probe, GPU suite, latency sweep per graph mode, headline
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe scripts/graph_while_probe.cu && /tmp/probe
( time KBA_GRAPH_VERBOSE=1 timeout 900 python -m pytest tests -x -q -m gpu --durations=3 ) > gpurun_out/r2_t20.log 2>&1; echo "pytest rc=$?"; grep -c "solve graph built" gpurun_out/r2_t20.log; grep "not available" gpurun_out/r2_t20.log | sort | uniq -c | head -3; tail -8 gpurun_out/r2_t20.log | cut -c1-200
for m in 0 1 2; do KBA_GRAPH=$m KBA_GRAPH_VERBOSE=1 timeout 300 python scripts/latency_sweep.py 1 64 2>&1 | tail -4; done | tee gpurun_out/r2_latency20.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-sub > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench20.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench20.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f (seq %.1f) ms/step %.1f lin frac %.3f launch ms %.3f jac-alone frac %.3f launches %d" % (d["value"], d["e2e"]["value"], d["e2e"]["sequential_value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["launch_ms_mean"], d["roofline_jacobian_kernel"]["frac"], d["gpu_launches"]))
PY
KBA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches20.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu20.log 2>&1; echo "ncu rc=$?"
python scripts/summarise_launches.py gpurun_out/r2_launches20.csv 2>/dev/null | head -30

#!/bin/bash
# validation of the committed state: the driver's own sequence (full GPU suite, smoke, default bench) + reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=6 ) > gpurun_out/r2_t17.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_t17.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench17.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench17.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f (seq %.1f) ms/step %.1f lin frac %.3f jac-alone frac %.3f (%.4f ms) launches %d" % (d["value"], d["e2e"]["value"], d["e2e"]["sequential_value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline_jacobian_kernel"]["frac"], d["roofline_jacobian_kernel"]["launch_ms"], d["gpu_launches"]))
print(json.dumps(d["sub_records"]["config2_batch"])[:700])
print(json.dumps(d["clocks"]))
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 | cut -c1-600
for lb in 2 3; do
KBA_LIN_BLOCKS=$lb timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lin_blocks $lb: value %.1f e2e %.1f ms/step %.1f lin_ms %.4f frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_mean'], d['roofline']['frac']))"
done
KBA_LIN_BLOCKS=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config2 or config3 or one_kernel or edge" 2>&1 | tail -2

#!/bin/bash
# round-2 baseline: GPU tests, full bench line, launch list, ncu --set full of one pass
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t9.log
tail -8 gpurun_out/r2_t9.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench9.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench9.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f (seq %.1f) ms/step %.1f pack ms %.1f jac_ms %.4f frac %.3f share %.3f" % (d["value"], d["e2e"]["value"], d["e2e"]["sequential_value"], d["ms_per_step"], d["e2e"]["host_pack_upload_ms_per_step"], d["roofline"]["launch_ms_mean"], d["roofline"]["frac"], d["roofline"]["share_of_timed_region"]))
print(json.dumps(d["cpu_baseline"])[:600])
for k, v in d["sub_records"].items(): print(k, json.dumps(v)[:1500])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches9.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu9.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches9.csv
# one full pass (skip the first ~120 launches: upload/packing + first passes), every kernel once or twice
timeout 900 ncu --set full --clock-control none --import-source on -s 150 -c 40 -o gpurun_out/r2_full9 -f python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu9_full.log 2>&1
ls -la gpurun_out/

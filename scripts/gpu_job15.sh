#!/bin/bash
# parallel tile builder, quad-per-row panel substitution, lm_scale prefetch: tests, phase clocks, full bench line, launch lists
mkdir -p gpurun_out
timeout 300 tests/cpp/test_facade gpu > gpurun_out/r2_facade15.log 2>&1; echo "facade rc=$?"; tail -2 gpurun_out/r2_facade15.log
timeout 900 python -m pytest tests/test_track.py tests/test_gpu_parity.py -m gpu -q -k "not config5_full" > gpurun_out/r2_t15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t15.log
grep -n "AssertionError\|passed\|failed" gpurun_out/r2_t15.log | cut -c1-250 | head -20
KBA_LIB_PATH=$PWD/limo_b200/libkba_b200_prof.so timeout 300 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 1 --in-flight 1 2>&1 | grep "reduced_solve cycles" | sed -n '20,22p'
timeout 900 python bench.py > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench15.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench15.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f (seq %.1f) ms/step %.1f pack ms %.1f lin_ms %.4f frac %.3f share %.3f" % (d["value"], d["e2e"]["value"], d["e2e"]["sequential_value"], d["ms_per_step"], d["e2e"]["host_pack_upload_ms_per_step"], d["roofline"]["launch_ms_mean"], d["roofline"]["frac"], d["roofline"]["share_of_timed_region"]))
for k, v in d["sub_records"].items(): print(k, json.dumps(v)[:900])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches15.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu15.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches15.csv | grep -v "k_pack\|k_track"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches15_b1.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 1 --in-flight 1 > gpurun_out/r2_ncu15_b1.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches15_b1.csv | grep -v "k_pack\|k_track"

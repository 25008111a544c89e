#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6.log
tail -12 gpurun_out/r2_t6.log | cut -c1-300
for dp in 1 0; do
KBA_DEVICE_PACK=$dp timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-sub > gpurun_out/r2_b6_dp$dp.json 2> gpurun_out/r2_b6_dp$dp.err; echo "bench device_pack=$dp rc=$?"; tail -3 gpurun_out/r2_b6_dp$dp.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_b6_dp$dp.json").read().strip().splitlines()[-1])
    print("device_pack=$dp value %.1f e2e %.1f (seq %.1f) ms/step %.1f pack+upload ms %.1f dl ms %.1f h2d %d jac_ms %.4f frac %.3f" % (d["value"], d["e2e"]["value"], d["e2e"]["sequential_value"], d["ms_per_step"], d["e2e"]["host_pack_upload_ms_per_step"], d["e2e"]["download_ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["roofline"]["launch_ms_mean"], d["roofline"]["frac"]))
except Exception as e:
    print("failed", e)
PY
done
KBA_LIB_PATH=$PWD/limo_b200/libkba_b200_prof.so timeout 300 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 2>&1 | grep "kba prof" | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches6.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu6.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches6.csv

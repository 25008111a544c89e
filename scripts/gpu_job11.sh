#!/bin/bash
# fused linearisation (k_linearize): correctness subset, A/B bench, launch lists (batch 148 and 1), ncu --set full, phase clocks
mkdir -p gpurun_out
( time timeout 300 tests/cpp/test_facade gpu ) > gpurun_out/r2_facade11.log 2>&1; echo "facade rc=$?"; tail -12 gpurun_out/r2_facade11.log
timeout 900 python -m pytest tests/test_track.py tests/test_gpu_parity.py -m gpu -q --durations=25 -k "not config5_full and not fp32" > gpurun_out/r2_t11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t11.log
tail -45 gpurun_out/r2_t11.log | cut -c1-300
for lz in 1 0; do
KBA_LINEARIZE=$lz timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('linearize $lz: value %.1f e2e %.1f ms/step %.1f lin_ms %.4f frac %.3f share %.3f iters %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_mean'], d['roofline']['frac'], d['roofline']['share_of_timed_region'], d['config']['lm_iterations_per_window_mean']))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches11.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu11.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches11.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches11_b1.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 1 --in-flight 1 > gpurun_out/r2_ncu11_b1.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches11_b1.csv
KBA_LIB_PATH=$PWD/limo_b200/libkba_b200_prof.so timeout 300 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 1 --in-flight 1 2>&1 | grep "reduced_solve cycles" | sed -n '20,24p'
timeout 900 ncu --set full --clock-control none --import-source on -s 140 -c 24 -o gpurun_out/r2_full11 -f python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu11_full.log 2>&1
ls -la gpurun_out/ | tail -5

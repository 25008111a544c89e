#!/bin/bash
# k_linearize v4 (cost at iteration zero only, division-free damping, lane-local segment sums), window-independent Schur split
mkdir -p gpurun_out
( time timeout 300 tests/cpp/test_facade gpu ) > gpurun_out/r2_facade14.log 2>&1; echo "facade rc=$?"; tail -4 gpurun_out/r2_facade14.log
timeout 900 python -m pytest tests/test_track.py tests/test_gpu_parity.py -m gpu -q -k "not config5_full and not fp32" > gpurun_out/r2_t14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t14.log
grep -n "AssertionError\|passed\|failed" gpurun_out/r2_t14.log | cut -c1-250 | head -20
for lz in 1 0; do
KBA_LINEARIZE=$lz timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('linearize $lz: value %.1f e2e %.1f ms/step %.1f lin_ms %.4f frac %.3f share %.3f iters %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_mean'], d['roofline']['frac'], d['roofline']['share_of_timed_region'], d['config']['lm_iterations_per_window_mean']))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches14.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu14.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches14.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches14_b1.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 1 --in-flight 1 > gpurun_out/r2_ncu14_b1.log 2>&1
python scripts/summarise_launches.py gpurun_out/r2_launches14_b1.csv | grep -v "k_pack\|k_track"
timeout 900 ncu --set full --clock-control none --import-source on -s 140 -c 24 -o gpurun_out/r2_full14 -f python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-sub --batch 148 --in-flight 1 > gpurun_out/r2_ncu14_full.log 2>&1
ls -la gpurun_out/ | tail -3

#!/bin/bash
# A/B on the headline workload: v0 = previous commit, v1 = all changes, v2 = v1 without the descriptor-only tiles of k_linearize,
# v3 = v2 with the previous k_backsub_v; v2 also with the WHILE-node graph
mkdir -p gpurun_out
python scripts/ab_variants.py ab_tmp/libkba_v0.so ab_tmp/libkba_v1.so ab_tmp/libkba_v2.so ab_tmp/libkba_v3.so ab_tmp/libkba_v2.so:2 ab_tmp/libkba_v0.so ab_tmp/libkba_v2.so 2>&1 | tee gpurun_out/r2_ab21.log

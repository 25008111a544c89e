#!/bin/bash
# per-window CTA loops: CTAs of k_linearize / k_backsub_v stride over the window's units (KBA_LIN_GRID / KBA_BS_GRID), A/B on the
# headline workload with a digest of the results
mkdir -p gpurun_out
L=limo_b200/libkba_b200.so
python scripts/ab_variants.py $L $L:0:KBA_LIN_GRID=96 $L:0:KBA_BS_GRID=64 $L:0:KBA_LIN_GRID=96,KBA_BS_GRID=64 $L:2:KBA_LIN_GRID=96,KBA_BS_GRID=64 $L:2 2>&1 | tee gpurun_out/r2_ab24.log

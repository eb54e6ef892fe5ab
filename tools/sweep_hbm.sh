#!/bin/bash
# Tuning sweep of the HBM-bound kernels at the B=128 shapes (tools/norm_bench.py): grid cap of the grid-stride kernels
# (SD2_WAVES_PCT) and the GroupNorm cluster configuration (SD2_GN_P0_KB / SD2_GN_P0_THREADS).  Output: gpurun_out/sweep_hbm.md
mkdir -p gpurun_out
out=gpurun_out/sweep_hbm.md
: > $out
for pct in 50 100 200 400; do
  echo "== SD2_WAVES_PCT=$pct" >> $out
  for f in geglu colsum ln_; do
    NORM_BENCH_BIG=1 SD2_WAVES_PCT=$pct timeout 300 python tools/norm_bench.py $f >> $out 2>&1
  done
done
for cfg in "110 256" "72 256" "225 512"; do
  set -- $cfg
  echo "== SD2_GN_P0_KB=$1 SD2_GN_P0_THREADS=$2" >> $out
  NORM_BENCH_BIG=1 SD2_GN_P0_KB=$1 SD2_GN_P0_THREADS=$2 timeout 300 python tools/norm_bench.py gn_ >> $out 2>&1
done
cat $out

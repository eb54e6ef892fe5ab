#!/bin/bash
# Runs the per-kernel GPU parity tests group by group in separate processes (a trapped kernel poisons the CUDA
# context of its own process only) and collects the logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in k1 linear conv3x3 attention groupnorm layernorm "geglu or layout or colsum or casts"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --maxfail=8 -k "$grp" > "gpurun_out/kt_${name}.log" 2>&1
  echo "== $grp: exit $? ==" | tee -a gpurun_out/kt_summary.txt
  tail -n 25 "gpurun_out/kt_${name}.log"
done

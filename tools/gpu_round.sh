#!/bin/bash
# One GPU-box session: GPU parity tests, the default bench line, (optionally) the GEMM autotune, then the ncu passes of
# the SAME bench command (launch list with DRAM bytes over one profiled step; one --set full capture of the top GEMM).
# Everything lands under gpurun_out/.  Usage: tools/gpu_round.sh [tests] [bench] [tune] [ncu] [ncufull] [b512] [parity]
mkdir -p gpurun_out
want() { [[ " $ARGS " == *" $1 "* ]]; }
ARGS="$*"
[ -z "$ARGS" ] && ARGS="tests bench ncu"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
if want tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
  echo "tests exit $?" | tee -a gpurun_out/summary.txt; tail -n 5 gpurun_out/tests_gpu.log
fi
if want smoke; then
  timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/smoke.log
fi
if want bench; then
  timeout 900 python bench.py > gpurun_out/bench_256.log 2> gpurun_out/bench_256.err
  echo "bench exit $?" | tee -a gpurun_out/summary.txt; tail -n 2 gpurun_out/bench_256.log; tail -n 5 gpurun_out/bench_256.err
fi
if want tune; then
  timeout 1200 python tools/autotune_gemm.py 128 32 gpurun_out/plans_128_32.json > gpurun_out/tune_128_32.log 2>&1
  echo "tune exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/tune_128_32.log
  if [ -s gpurun_out/plans_128_32.json ]; then
    python - <<'E'
import json, os
p = 'diffusion_b200/gemm_plans.json'
t = json.load(open(p)) if os.path.exists(p) else {}
t.update(json.load(open('gpurun_out/plans_128_32.json')))
json.dump(t, open(p, 'w'), indent=0, sort_keys=True)
E
    timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_256_tuned.log 2> gpurun_out/bench_256_tuned.err
    echo "bench tuned exit $?" | tee -a gpurun_out/summary.txt; tail -n 2 gpurun_out/bench_256_tuned.log
  fi
fi
if want tune512; then
  timeout 1200 python tools/autotune_gemm.py 32 64 gpurun_out/plans_32_64.json > gpurun_out/tune_32_64.log 2>&1
  echo "tune512 exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/tune_32_64.log
fi
if want b256batch; then
  timeout 900 python bench.py --batch 256 --steps 10 --no-cpu-baseline > gpurun_out/bench_256_B256.log 2> gpurun_out/bench_256_B256.err
  echo "bench B256 exit $?" | tee -a gpurun_out/summary.txt; tail -n 2 gpurun_out/bench_256_B256.log; tail -n 3 gpurun_out/bench_256_B256.err
fi
if want normsweep; then
  for cfg in "110 256" "72 256" "56 256" "110 512" "225 512"; do
    set -- $cfg
    echo "== SD2_GN_P0_KB=$1 SD2_GN_P0_THREADS=$2" >> gpurun_out/norm_sweep.md
    NORM_BENCH_BIG=1 SD2_GN_P0_KB=$1 SD2_GN_P0_THREADS=$2 timeout 300 python tools/norm_bench.py gn_ >> gpurun_out/norm_sweep.md 2>&1
  done
  echo "== other HBM kernels (B=128 shapes)" >> gpurun_out/norm_sweep.md
  NORM_BENCH_BIG=1 timeout 300 python tools/norm_bench.py l >> gpurun_out/norm_sweep.md 2>&1
  echo "normsweep exit $?" | tee -a gpurun_out/summary.txt; tail -n 12 gpurun_out/norm_sweep.md
fi
if want b512big; then
  timeout 900 python bench.py --latent 64 --batch 64 --steps 5 --no-cpu-baseline > gpurun_out/bench_512_B64.log 2> gpurun_out/bench_512_B64.err
  echo "bench512 B64 exit $?" | tee -a gpurun_out/summary.txt; tail -n 2 gpurun_out/bench_512_B64.log; tail -n 3 gpurun_out/bench_512_B64.err
fi
if want b512; then
  timeout 900 python bench.py --latent 64 --steps 10 --no-cpu-baseline > gpurun_out/bench_512.log 2> gpurun_out/bench_512.err
  echo "bench512 exit $?" | tee -a gpurun_out/summary.txt; tail -n 2 gpurun_out/bench_512.log
fi
if want parity; then
  timeout 900 python tools/step_parity.py sd2 8 32 > gpurun_out/parity_sd2_B8.log 2>&1
  echo "parity exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/parity_sd2_B8.log
fi
if want ncu; then
  timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-step > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
fi
if want ncufull; then
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel \
    --launch-skip 200 -c 3 -o gpurun_out/gemm_full -f python bench.py --profile-step > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?" | tee -a gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt

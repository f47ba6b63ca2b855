#!/bin/bash
# ncu --set full captures of the 'next'-row kernels (each only after the same command exited 0 without ncu).
TAG=${1:-next}
mkdir -p gpurun_out
for pair in "midas_mse:midas_loss_kernel" "robust:robust_stats_kernel" "stdepth:stdepth_loss_kernel"; do
  n=${pair%%:*}; k=${pair##*:}
  python tools/run_one.py $n 3 > gpurun_out/plain_${n}_$TAG.log 2>&1 || { echo "plain $n failed"; tail -3 gpurun_out/plain_${n}_$TAG.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 2 -c 1 -f -o gpurun_out/prof_${n}_$TAG python tools/run_one.py $n 3 > gpurun_out/ncu_${n}_$TAG.log 2>&1
  echo "$n rc=$?"
done

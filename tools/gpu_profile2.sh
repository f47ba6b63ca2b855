#!/bin/bash
# ncu --set full captures of the non-headline kernels (one launch each, after a plain run exited 0)
mkdir -p gpurun_out
for n in c5_metrics dorn_fused; do
  python tools/run_one.py $n 3 > gpurun_out/plain_$n.log 2>&1 || { echo "plain $n failed"; tail -3 gpurun_out/plain_$n.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:"metrics_kernel|dorn_kernel|vnl_kernel|masked_loss_kernel|ord_loss_kernel" -s 1 -c 1 -f -o gpurun_out/prof_$n python tools/run_one.py $n 3 > gpurun_out/ncu_$n.log 2>&1
  echo "$n rc=$?"
done

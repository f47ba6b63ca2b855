#!/bin/bash
# ncu evidence for the bench command and the big-config kernels. Every ncu run only after the SAME command
# exited 0 without ncu. Outputs under gpurun_out/ (launch list CSV, .ncu-rep files); summarised by
# tools/ncu_summary.py into profiles/.   usage: tools/gpu_profile_r1.sh [tag]
TAG=${1:-a}
mkdir -p gpurun_out
CMD="python bench.py --steps 16 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
for n in c2_fused c5_metrics dorn_fused; do
  python tools/run_one.py $n 3 > gpurun_out/plain_${n}_$TAG.log 2>&1 || { echo "plain $n failed"; tail -3 gpurun_out/plain_${n}_$TAG.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:"metrics_kernel|dorn_kernel|masked_loss_kernel" -s 2 -c 1 -f -o gpurun_out/prof_${n}_$TAG python tools/run_one.py $n 3 > gpurun_out/ncu_${n}_$TAG.log 2>&1
  echo "$n rc=$?"
done
ls -la gpurun_out/ | tail -20

#!/bin/bash
# ncu evidence for the bench command: (1) launch list with per-launch device time, (2) one --set full
# capture of the two hot kernels. Each ncu run only after the same command exited 0 without ncu.
mkdir -p gpurun_out
CMD="python bench.py --steps 16 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"masked_loss_kernel|metrics_kernel" -s 12 -c 4 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/

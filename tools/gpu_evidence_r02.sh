#!/bin/bash
# Round-2 evidence on one box: ncu launch list of the bench command and one --set full capture per hot kernel (each only
# after the same command exited 0 without ncu). Files land in gpurun_out/; tools/ncu_summary.py turns them into profiles/.
mkdir -p gpurun_out
CMD="python bench.py --steps 16 --warmup 3 --no-graph --no-cpu-baseline --no-configs --no-eager-gpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
for n in ${EVID_KERNELS:-c2_fused c5_metrics c5_metrics10 stdepth dorn_fused}; do
  python tools/run_one.py $n 3 > gpurun_out/plain_$n.log 2>&1 || { echo "plain $n failed"; tail -3 gpurun_out/plain_$n.log; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"silog_ss_kernel|metrics_kernel|stdepth|dorn_kernel" -s 2 -c 1 -f -o gpurun_out/prof_r02_$n python tools/run_one.py $n 3 > gpurun_out/ncu_$n.log 2>&1
  echo "$n rc=$?"
done

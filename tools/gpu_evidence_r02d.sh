#!/bin/bash
# Final evidence of round 2, session 4, on one box: the whole GPU suite in one process (as the driver runs it), smoke, the default
# bench line, the ncu launch list of the bench command and --set full captures of the kernels this session changed (DORN
# supervision step and decode at C3, register-resident MaskedDepthLoss at C1), each ncu pass only after the same command exited 0
# without ncu. Files land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02d_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
CMD="python bench.py --steps 16 --warmup 3 --no-graph --no-cpu-baseline --no-configs --no-eager-gpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
for n in dorn_fused:dorn_kernel dorn_decode:dorn_kernel c1_eigen:eigen_resident_kernel; do
  name=${n%%:*}; kern=${n##*:}
  python tools/run_one.py $name 3 > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name failed"; tail -3 gpurun_out/plain_$name.log; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$kern" -s 2 -c 1 -f -o gpurun_out/prof_r02d_$name python tools/run_one.py $name 3 > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
done

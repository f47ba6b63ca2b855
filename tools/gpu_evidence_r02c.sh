#!/bin/bash
# Final evidence of round 2, session 3, on one box: every GPU test file, smoke, the default bench line, the ncu launch list of
# the bench command and one --set full capture of the register-resident small-input kernel (berHu + 7 metrics at C1), each
# ncu pass only after the same command exited 0 without ncu. Files land in gpurun_out/.
mkdir -p gpurun_out
BENCH_STEPS=${BENCH_STEPS:-100} bash tools/gpu_round.sh
echo "gpu_round rc=$?"
CMD="python bench.py --steps 16 --warmup 3 --no-graph --no-cpu-baseline --no-configs --no-eager-gpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
python tools/run_one.py c1_berhu 3 > gpurun_out/plain_c1_berhu.log 2>&1 || { echo "plain c1_berhu failed"; tail -3 gpurun_out/plain_c1_berhu.log; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"resident_loss_kernel" -s 2 -c 1 -f -o gpurun_out/prof_r02c_c1_berhu python tools/run_one.py c1_berhu 3 > gpurun_out/ncu_c1_berhu.log 2>&1
echo "c1_berhu rc=$?"
timeout 120 python tools/c1_probe.py > gpurun_out/c1_probe_final.jsonl 2> gpurun_out/c1_probe_final.err
MDE_NO_RESIDENT=1 timeout 120 python tools/c1_probe.py > gpurun_out/c1_probe_final_generic.jsonl 2> gpurun_out/c1_probe_final_generic.err
tail -n 5 gpurun_out/c1_probe_final.jsonl

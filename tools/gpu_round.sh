#!/bin/bash
# One GPU-box round: parity tests per file (a faulting kernel must not poison the other files),
# smoke, bench. Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$n.log 2>&1 || rc=1
  tail -n 3 gpurun_out/$n.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 || rc=1
tail -n 2 gpurun_out/smoke.log
timeout 400 python bench.py --steps ${BENCH_STEPS:-100} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err || rc=1
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
exit $rc

#!/bin/bash
# A/B of SS-kernel variants on one box: tools/gpu_ab.sh v1 v2 ... (tools/variants/<v>.so), then the phase trace of the
# instrumented twin, the loss / config-size / midas parity files and a short bench.
mkdir -p gpurun_out
for v in "$@"; do
  MDE_B200_LIB=$PWD/tools/variants/$v.so timeout 300 python tools/check_c2.py > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo "$v: $(cat gpurun_out/ab_$v.json | python -c 'import json,sys; d=json.load(sys.stdin); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.endswith("_us")}, {k:(v.get("loss_rel"),v.get("grad_rel_max"),v.get("metric_rel_max"),v.get("counts_match")) for k,v in d.items() if isinstance(v,dict)})' 2>&1)"
done
# second pass in reverse order (clock / thermal drift check)
for v in $(echo "$@" | tr ' ' '\n' | tac); do
  MDE_B200_LIB=$PWD/tools/variants/$v.so timeout 300 python tools/check_c2.py --no-parity > gpurun_out/ab2_$v.json 2> gpurun_out/ab2_$v.err
  echo "$v (2): $(cat gpurun_out/ab2_$v.json)"
done
if [ -f tools/libmde_dbg.so ]; then
  MDE_B200_LIB=$PWD/tools/libmde_dbg.so MDE_TRACE_FUSED=1 timeout 300 python tools/trace_dump.py > gpurun_out/trace_fused.log 2>&1
  MDE_B200_LIB=$PWD/tools/libmde_dbg.so MDE_TRACE_FUSED=0 timeout 300 python tools/trace_dump.py > gpurun_out/trace_plain.log 2>&1
fi
for f in ${AB_TESTS:-tests/test_gpu_losses.py tests/test_gpu_config_size.py tests/test_gpu_midas.py}; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  tail -n 3 gpurun_out/$n.log
done
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu ${AB_BENCH_FLAGS:---no-configs} > gpurun_out/b.json 2> gpurun_out/b.err
cat gpurun_out/b.json; tail -n 3 gpurun_out/b.err

# DORN head: parity tests of the default build, then the C3 probe (tools/dorn_probe.py) per library on the same box
timeout 300 python -m pytest tests/test_gpu_dorn.py -x -q > gpurun_out/t_dorn.log 2>&1; echo rc=$? >> gpurun_out/t_dorn.log; tail -5 gpurun_out/t_dorn.log
: > gpurun_out/dorn_ab3.jsonl
for v in dorn_head ""; do
  if [ -z "$v" ]; then timeout 120 python tools/dorn_probe.py >> gpurun_out/dorn_ab3.jsonl 2>> gpurun_out/dorn_ab.err
  else MDE_B200_LIB=$PWD/tools/variants/$v.so timeout 120 python tools/dorn_probe.py >> gpurun_out/dorn_ab3.jsonl 2>> gpurun_out/dorn_ab.err; fi
done
cat gpurun_out/dorn_ab3.jsonl; tail -3 gpurun_out/dorn_ab.err

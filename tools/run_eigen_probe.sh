set -x
timeout 300 python -m pytest tests/test_gpu_resident.py -x -q -k eigen > gpurun_out/t_eigen_res.log 2>&1; echo rc=$? >> gpurun_out/t_eigen_res.log
tail -15 gpurun_out/t_eigen_res.log
timeout 120 python tools/eigen_probe.py > gpurun_out/eigen_probe.jsonl 2> gpurun_out/eigen_probe.err; cat gpurun_out/eigen_probe.jsonl; tail -3 gpurun_out/eigen_probe.err
MDE_NO_RESIDENT=1 timeout 120 python tools/eigen_probe.py > gpurun_out/eigen_probe_generic.jsonl 2>/dev/null; cat gpurun_out/eigen_probe_generic.jsonl

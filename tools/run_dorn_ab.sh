# DORN head: parity tests of the default build, then the C3 probe (tools/dorn_probe.py); MDE_B200_LIB selects another build for A/B
timeout 300 python -m pytest tests/test_gpu_dorn.py tests/test_gpu_config_size.py -x -q > gpurun_out/t_dorn.log 2>&1; echo rc=$? >> gpurun_out/t_dorn.log; tail -3 gpurun_out/t_dorn.log
timeout 120 python tools/dorn_probe.py > gpurun_out/dorn_final.jsonl 2>> gpurun_out/dorn_ab.err; cat gpurun_out/dorn_final.jsonl

timeout 600 python -m pytest tests/test_gpu_dropin.py -q -x 2>&1 | tail -25 > gpurun_out/r2_t15.log; tail -25 gpurun_out/r2_t15.log

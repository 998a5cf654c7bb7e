#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_plan_loop.py tests/test_gpu_parity.py tests/test_gpu_target.py -q -m gpu -x > gpurun_out/r2o_pytest.txt 2>&1
echo "rc=$?" >> gpurun_out/r2o_pytest.txt
tail -n 40 gpurun_out/r2o_pytest.txt | cut -c1-250
python tools/step_latency.py c2 > gpurun_out/r2_step_latency2.txt 2>&1; python tools/step_latency.py c1 >> gpurun_out/r2_step_latency2.txt 2>&1; grep "ms per" gpurun_out/r2_step_latency2.txt

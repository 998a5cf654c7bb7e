#!/bin/bash
# quick A/B of a fused-eval change: parity suites that touch the fused gradient, then the eval-only bench at 1e7 and at the 8-GPU slice
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_pairforms.py tests/test_gpu_exchange.py tests/test_gpu_plan_loop.py -q -m gpu -x > gpurun_out/r2j_pytest.txt 2>&1
echo "rc=$?" >> gpurun_out/r2j_pytest.txt
tail -n 12 gpurun_out/r2j_pytest.txt | cut -c1-250
for N in 10000000 1250000; do
timeout 600 python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu --no-also --samples $N > gpurun_out/r2j_bench_$N.json 2> gpurun_out/r2j_bench_$N.err
tail -n 2 gpurun_out/r2j_bench_$N.err
python -c "import json;d=json.loads(open('gpurun_out/r2j_bench_$N.json').read().strip().splitlines()[-1]);print($N, d['ms_per_step'],d['roofline']['frac'])"
done

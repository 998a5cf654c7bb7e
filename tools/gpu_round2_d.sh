#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2j_pytest_gpu.txt 2>&1
echo "suite rc=$?" >> gpurun_out/r2j_pytest_gpu.txt
tail -n 25 gpurun_out/r2j_pytest_gpu.txt | cut -c1-300
timeout 900 python bench.py --steps 200 --warmup 5 --no-cpu --no-also --e2e-steps 5 > gpurun_out/r2j_bench_c4.json 2> gpurun_out/r2j_bench_c4.err
tail -n 3 gpurun_out/r2j_bench_c4.err
python - gpurun_out/r2j_bench_c4.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"])
print("e2e", d["e2e"])
PY
python - <<'PY'
import sys, time
sys.path[:0]=['.','embodied-active-learning-vision_b200']
import torch
from control_torch import engine
low=torch.tensor([-1.15]*6); high=torch.tensor([1.15]*6)
for n in (100000, 1250000, 10000000):
    torch.manual_seed(0)
    engine.device_uniform(1000, low, high)
    torch.cuda.synchronize(); t0=time.perf_counter()
    x=engine.device_uniform(n, low, high)
    torch.cuda.synchronize(); t1=time.perf_counter()
    print("device_uniform", n, "rows x 6:", (t1-t0)*1e3, "ms")
PY

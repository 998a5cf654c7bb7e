#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2l_pytest_gpu.txt 2>&1
echo "suite rc=$?" >> gpurun_out/r2l_pytest_gpu.txt
tail -n 40 gpurun_out/r2l_pytest_gpu.txt | cut -c1-250
timeout 1200 python bench.py --steps 200 --no-e2e --no-cpu > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -n 5 gpurun_out/r2l_bench.err
python - gpurun_out/r2l_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"])
for k,v in d["also"].items():
    print(k, {kk:vv for kk,vv in v.items() if not isinstance(vv,(dict,str))})
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2k_pytest_gpu.txt 2>&1
echo "suite rc=$?" >> gpurun_out/r2k_pytest_gpu.txt
tail -n 6 gpurun_out/r2k_pytest_gpu.txt | cut -c1-300
( time timeout 1200 python bench.py > gpurun_out/r2k_bench_default.json 2> gpurun_out/r2k_bench_default.err ) 2>&1 | tail -3
tail -n 5 gpurun_out/r2k_bench_default.err
python - gpurun_out/r2k_bench_default.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["config"]["note"])
e=d["e2e"]; print("e2e ms", e["ms_per_robot_step"], "host_draw", e.get("host_draw"), e["pair_mix"], e["h2d_bytes_per_step"])
print("cpu", d["cpu_baseline"])
for k,v in d["also"].items():
    print(k, {kk:vv for kk,vv in v.items() if not isinstance(vv,(dict,str))})
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err ) 2>&1 | tail -3
cat gpurun_out/r2k_bench_ref.json | cut -c1-900

#!/bin/bash
# what the driver runs at round end: GPU test suite, smoke, default bench (both arms)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2_final_pytest_gpu.txt 2>&1
echo "suite rc=$?" >> gpurun_out/r2_final_pytest_gpu.txt
tail -n 25 gpurun_out/r2_final_pytest_gpu.txt | cut -c1-250
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.txt 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_final_smoke.txt
tail -n 8 gpurun_out/r2_final_smoke.txt | cut -c1-250
timeout 1500 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "bench rc=$?"
tail -n 5 gpurun_out/r2_final_bench.err
python - gpurun_out/r2_final_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
print("e2e", {k:v for k,v in d["e2e"].items() if not isinstance(v,(dict,str))})
print("e2e extras", d["e2e"].get("host_draw"), d["e2e"].get("also"))
print("cpu", d["cpu_baseline"])
for k,v in d["also"].items():
    print(k, {kk:vv for kk,vv in v.items() if not isinstance(vv,(dict,str))})
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
echo "ref rc=$?"; cut -c1-600 gpurun_out/r2_final_bench_ref.json

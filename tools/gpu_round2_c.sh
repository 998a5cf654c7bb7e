#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2h_pytest_gpu.txt 2>&1
echo "suite rc=$?" >> gpurun_out/r2h_pytest_gpu.txt
tail -n 40 gpurun_out/r2h_pytest_gpu.txt | cut -c1-300
for n in 10000000 1250000; do
  KLERG_VARIANT=_stamps timeout 300 python tools/cta_timeline.py c4 $n > gpurun_out/r2h_timeline_c4_$n.txt 2>&1
done
cat gpurun_out/r2h_timeline_*.txt
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu --no-also --e2e-steps 3 > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err
KLERG_MIXED_WARPS=16 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --no-also > gpurun_out/r2h_bench_c4_w16.json 2> gpurun_out/r2h_bench_c4_w16.err
timeout 600 python bench.py --steps 1000 --warmup 5 --no-cpu --no-e2e --no-also --samples 1250000 > gpurun_out/r2h_bench_c4_1p25M.json 2> gpurun_out/r2h_bench_c4_1p25M.err
timeout 600 python bench.py --steps 2000 --warmup 5 --no-cpu --no-e2e --no-also --workload c2 > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err
for f in gpurun_out/r2h_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"] if d.get("roofline") else None, d["config"]["note"])
    if d.get("e2e"): print("e2e", d["e2e"]["ms_per_robot_step"], d["e2e"]["value"])
except Exception as e:
    print("ERR", e)
PY
done
tail -n 5 gpurun_out/r2h_bench_c4.err

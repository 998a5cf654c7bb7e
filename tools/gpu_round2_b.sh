#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exchange.py -q -m gpu > gpurun_out/r2b_pytest_exchange.txt 2>&1
echo "exchange rc=$?" >> gpurun_out/r2b_pytest_exchange.txt
for n in 10000000 1250000; do
  KLERG_VARIANT=_stamps timeout 300 python tools/cta_timeline.py c4 $n > gpurun_out/r2b_timeline_c4_$n.txt 2>&1
done
KLERG_VARIANT=_stamps timeout 300 python tools/cta_timeline.py c2 100000 > gpurun_out/r2b_timeline_c2.txt 2>&1
KLERG_PDL=0 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --no-also > gpurun_out/r2b_bench_c4_nopdl.json 2> gpurun_out/r2b_bench_c4_nopdl.err
KLERG_PDL=0 timeout 600 python bench.py --steps 1000 --warmup 5 --no-cpu --no-e2e --no-also --samples 1250000 > gpurun_out/r2b_bench_c4_1p25M_nopdl.json 2> gpurun_out/r2b_bench_c4_1p25M_nopdl.err
timeout 600 python bench.py --steps 1000 --warmup 5 --no-cpu --no-e2e --no-also --samples 1250000 --no-graph > gpurun_out/r2b_bench_c4_1p25M_nograph.json 2> gpurun_out/r2b_bench_c4_1p25M_nograph.err
tail -n 5 gpurun_out/r2b_pytest_exchange.txt
cat gpurun_out/r2b_timeline_*.txt
for f in gpurun_out/r2b_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"] if d.get("roofline") else None, d["config"]["note"])
except Exception as e:
    print("ERR", e)
PY
done

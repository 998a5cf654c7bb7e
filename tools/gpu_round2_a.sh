#!/bin/bash
# first GPU pass of round 2: exchange-protocol tests, the whole GPU suite, quick benches (1 GPU)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2a_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_exchange.py -x -q -m gpu > gpurun_out/r2a_pytest_exchange.txt 2>&1
echo "exchange rc=$?" >> gpurun_out/r2a_pytest_exchange.txt
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_exchange.py > gpurun_out/r2a_pytest_gpu.txt 2>&1
echo "suite rc=$?" >> gpurun_out/r2a_pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --no-also > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --no-also --no-overlap > gpurun_out/r2a_bench_c4_nooverlap.json 2> gpurun_out/r2a_bench_c4_nooverlap.err
timeout 600 python bench.py --steps 1000 --warmup 5 --no-cpu --no-e2e --no-also --samples 1250000 > gpurun_out/r2a_bench_c4_1p25M.json 2> gpurun_out/r2a_bench_c4_1p25M.err
timeout 600 python bench.py --steps 1000 --warmup 5 --no-cpu --no-e2e --no-also --samples 1250000 --no-overlap > gpurun_out/r2a_bench_c4_1p25M_nooverlap.json 2> gpurun_out/r2a_bench_c4_1p25M_nooverlap.err
timeout 600 python bench.py --steps 2000 --warmup 5 --no-cpu --no-e2e --no-also --workload c2 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
tail -3 gpurun_out/r2a_pytest_exchange.txt gpurun_out/r2a_pytest_gpu.txt
for f in gpurun_out/r2a_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"] if d.get("roofline") else None, d["config"]["note"])
except Exception as e:
    print("ERR", e)
PY
done

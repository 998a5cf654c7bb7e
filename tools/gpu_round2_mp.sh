#!/bin/bash
# multi-GPU pass: sharded parity tests + bench at the box's GPU count
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharding.py -q -m gpu > gpurun_out/r2_mp${N}_pytest.txt 2>&1
echo "rc=$?" >> gpurun_out/r2_mp${N}_pytest.txt
tail -n 15 gpurun_out/r2_mp${N}_pytest.txt | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 5 --no-cpu --e2e-steps 10 > gpurun_out/r2_mp${N}_bench.json 2> gpurun_out/r2_mp${N}_bench.err
echo "bench rc=$?"
tail -n 5 gpurun_out/r2_mp${N}_bench.err
KLERG_VARIANT=_stamps timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/cta_timeline.py c4 > gpurun_out/r2_mp${N}_timeline.txt 2>&1
tail -n 32 gpurun_out/r2_mp${N}_timeline.txt | cut -c1-220
python - gpurun_out/r2_mp${N}_bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches","n_gpus")}, d["config"]["note"])
    if d.get("e2e"): print("e2e ms/step", d["e2e"]["ms_per_robot_step"], "steps", d["e2e"]["steps"])
    print("rank_identical", d.get("rank_identical"), "roofline frac", d["roofline"]["frac"])
    if d.get("also"): print({k:(v.get("us_per_eval")) for k,v in d["also"].items() if isinstance(v,dict)})
except Exception as e:
    print("ERR", e)
PY

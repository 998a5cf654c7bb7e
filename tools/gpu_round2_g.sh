#!/bin/bash
# fingerprint belief update + batch kernel after the thread-count change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fingerprint.py tests/test_gpu_configs.py -q -m gpu -x > gpurun_out/r2m_pytest.txt 2>&1
echo "rc=$?" >> gpurun_out/r2m_pytest.txt
tail -n 30 gpurun_out/r2m_pytest.txt | cut -c1-250
timeout 900 python bench.py --steps 100 --no-e2e --no-cpu > gpurun_out/r2m_bench_c3.json 2> gpurun_out/r2m_bench_c3.err
tail -n 3 gpurun_out/r2m_bench_c3.err
python -c "import json;d=json.loads(open(\"gpurun_out/r2m_bench_c3.json\").read().strip().splitlines()[-1]);print(d[\"ms_per_step\"],d[\"roofline\"][\"frac\"]);[print(k,{a:b for a,b in v.items() if not isinstance(b,(dict,str))}) for k,v in d[\"also\"].items()]"
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, "embodied-active-learning-vision_b200")
from dist_modules.fingerprint_module import FingerprintDist
for states, lims in (("xy", [[-1, 1]] * 2), ("xyw", [[-1, 1], [-1, 1], [-2, 2]])):
    fd = FingerprintDist(explr_states=states, plot_idx=[0, 1], capacity=512, lims=[list(map(float, l)) for l in lims], name=("a", "b", "c"))
    rng = np.random.default_rng(0)
    for n in (1, 10, 100):
        ts = []
        for rep in range(5):
            fd.push_batch(rng.random((n, len(states))) - 0.5, rng.random(n))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            fd.update_prior()
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(f"belief_update {states} G={fd.grid.shape[0]} n={n}: {min(ts)*1e3:.3f} ms (host call incl. H2D of locs)")
PY

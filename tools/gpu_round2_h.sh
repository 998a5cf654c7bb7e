#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_pairforms.py tests/test_gpu_fingerprint.py -q -m gpu -x > gpurun_out/r2n_pytest.txt 2>&1
echo "rc=$?" >> gpurun_out/r2n_pytest.txt
tail -n 30 gpurun_out/r2n_pytest.txt | cut -c1-250
timeout 900 python bench.py --steps 100 --no-e2e --no-cpu > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -n 3 gpurun_out/r2n_bench.err
python -c "import json;d=json.loads(open(\"gpurun_out/r2n_bench.json\").read().strip().splitlines()[-1]);print(d[\"ms_per_step\"],d[\"roofline\"][\"frac\"]);[print(k,{a:b for a,b in v.items() if not isinstance(b,(dict,str))}) for k,v in d[\"also\"].items()]"

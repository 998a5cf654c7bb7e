#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2l_pytest.txt 2>&1
echo "rc=$?" >> gpurun_out/r2l_pytest.txt
tail -n 6 gpurun_out/r2l_pytest.txt | cut -c1-300
python tools/step_phases.py c2 c1 > gpurun_out/r2l_step_phases.txt 2>&1; grep "total\|prefetch\|device_uniform" gpurun_out/r2l_step_phases.txt
timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu --no-also > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -n 3 gpurun_out/r2l_bench.err
python -c "import json;d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['roofline']['frac']);print({k:v for k,v in d['e2e'].items() if not isinstance(v,(dict,str))}); print(d['e2e'].get('host_draw'))"

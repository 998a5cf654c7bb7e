#!/bin/bash
# a22 variants: whole GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2k_pytest.txt 2>&1
echo "rc=$?" >> gpurun_out/r2k_pytest.txt
tail -n 40 gpurun_out/r2k_pytest.txt | cut -c1-300

#!/bin/bash
# ncu --set full of the config-3 batch cost kernel (one launch), after a plain run of the same command
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/ncu_plain_batch.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:eval_cost_batch_kernel -c 1 -f -o gpurun_out/prof_r02_batch_c3 $CMD > gpurun_out/ncu_batch.log 2>&1
tail -n 5 gpurun_out/ncu_batch.log

#!/bin/bash
# ncu --set full of the fused eval kernel at c4 (one launch), after a plain run of the same command
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-also --profile-evals 2"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:eval_grad_kernel -c 1 -f -o gpurun_out/prof_r02_eval_c4 $CMD > gpurun_out/ncu_eval.log 2>&1
tail -n 5 gpurun_out/ncu_eval.log

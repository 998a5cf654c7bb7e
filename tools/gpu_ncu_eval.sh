#!/bin/bash
# ncu --set full of the fused eval kernel at c4 (one launch), after a plain run of the same command
# usage: gpu_ncu_eval.sh [samples]   (default: the workload's 1e7; 5000000 / 2500000 / 1250000 = the per-rank sizes at 2/4/8 GPUs)
mkdir -p gpurun_out
N=${1:-10000000}
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-also --samples $N --profile-evals 2"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:eval_grad_kernel -c 1 -f -o gpurun_out/prof_r02_eval_c4_$N $CMD > gpurun_out/ncu_eval.log 2>&1
tail -n 3 gpurun_out/ncu_eval.log
ncu -i gpurun_out/prof_r02_eval_c4_$N.ncu-rep --page raw --csv > gpurun_out/prof_r02_eval_c4_$N.raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/prof_r02_eval_c4_$N.raw.csv | head -24

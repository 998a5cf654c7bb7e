#!/bin/bash
# launch list (device time per launch) of the default bench command, after a plain run of the same command
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 2"
$CMD > gpurun_out/r2_launches_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench_c4.csv $CMD > gpurun_out/r2_launches_ncu.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/r2_launches_ncu.log | cut -c1-300
python profiles/summarize_launches.py gpurun_out/r02_launches_bench_c4.csv | head -40

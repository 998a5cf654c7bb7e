#!/bin/bash
# ncu --set full of the tensor-core history pass (one launch, 1e5 rows x 2e6 samples), after a plain run
mkdir -p gpurun_out
CMD="python tools/time_footprint.py c4 2000000"
$CMD > gpurun_out/ncu_ft_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:footprint_tc_kernel --launch-skip 2 -c 1 -f -o gpurun_out/prof_r02_footprint_tc $CMD > gpurun_out/ncu_ft.log 2>&1
tail -n 3 gpurun_out/ncu_ft.log
ncu -i gpurun_out/prof_r02_footprint_tc.ncu-rep --page raw --csv > gpurun_out/prof_r02_footprint_tc.raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/prof_r02_footprint_tc.raw.csv | head -30

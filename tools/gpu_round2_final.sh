#!/bin/bash
# final single-GPU evidence of the round: suite + smoke + default bench (both arms), ncu --set full of the fused eval,
# launch list, per-CTA timelines (instrumented build), step latency profile
bash tools/gpu_round2_full.sh
bash tools/gpu_ncu_eval.sh 10000000
bash tools/gpu_ncu_launches.sh
KLERG_VARIANT=_stamps timeout 300 python tools/cta_timeline.py c4 10000000 > gpurun_out/r2_timeline_1gpu_10000000.txt 2>&1
KLERG_VARIANT=_stamps timeout 300 python tools/cta_timeline.py c4 1250000 > gpurun_out/r2_timeline_1gpu_1250000.txt 2>&1
tail -n 16 gpurun_out/r2_timeline_1gpu_1250000.txt
python tools/step_latency.py c2 c1 > gpurun_out/r2_step_latency.txt 2>&1; grep "ms per" gpurun_out/r2_step_latency.txt
python tools/step_phases.py c2 c1 > gpurun_out/r2_step_phases.txt 2>&1; grep "total" gpurun_out/r2_step_phases.txt

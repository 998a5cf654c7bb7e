"""bench.py with the horizon of the workload overridden (KLERG_BENCH_H): how the per-pair rate of the fused eval depends
on the split of H over the 16 warps of the gradient pass (H = 48: 3 states each; 50: 14 x 3 + 2 x 4; 64: 4 each)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "embodied-active-learning-vision_b200"))
import workloads as wl
for k in ("c4",):
    wl.WORKLOADS[k]["H"] = int(os.environ["KLERG_BENCH_H"])
import bench
bench.main()

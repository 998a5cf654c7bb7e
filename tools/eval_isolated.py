"""Is the fused eval slower back-to-back than alone?  Times single c4 evals with CUDA events: isolated (sync + sleep
before each) against runs of 1, 2, 4, 16 evals in a row."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "embodied-active-learning-vision_b200"))
import torch
import workloads as wl
import bench
from control_torch import engine, _cabi as cabi
from control_torch.klerg import Robot
from control_torch.planner import PlannerContext
dev = torch.device("cuda")
S = bench.build_sets("c4", wl.WORKLOADS["c4"]["N"], 0, engine.SINGLE, dev, engine, Robot, PlannerContext, max_sets=2)
sets = S["sets"]
def ev(i):
    c = sets[i % len(sets)]
    return c.gradient(c.u)
for i in range(4):
    ev(i)
torch.cuda.synchronize()
def run(n, sleep):
    ts = []
    for rep in range(6):
        torch.cuda.synchronize()
        if sleep:
            time.sleep(sleep)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            ev(i)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / n * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]
for n, sl in ((1, 0.2), (1, 0.0), (2, 0.0), (4, 0.0), (16, 0.0), (64, 0.0), (1, 0.2)):
    lo, med = run(n, sl)
    print(f"{n:3d} evals in a row, sleep {sl:.1f} s before: {lo:7.1f} us min, {med:7.1f} us median per eval")

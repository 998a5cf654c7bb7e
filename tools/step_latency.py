#!/usr/bin/env python
"""Robot.step() latency at the reference's own operating sizes (configs 1 and 2), with a host profile."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "embodied-active-learning-vision_b200"))
import workloads as wl  # noqa: E402
from control_torch.klerg import Robot  # noqa: E402


def run(name, steps=50, profile=False):
    w = wl.WORKLOADS[name]
    lims = [wl.LIMS[s] for s in w["states"]]
    target = wl.make_target(w["target"], lims, seed=1, device=torch.device("cuda"))
    torch.manual_seed(7)
    r = Robot(process_group=None, **wl.robot_kwargs(name, target))
    r.test(1000)
    for row in wl.random_walk_history(name, w["M"], seed=5):
        r.memory_buffer.push(row)
    for _ in range(5):
        r.step(w["N"], w["M"], save_update=True)
    torch.cuda.synchronize()
    pr = cProfile.Profile() if profile else None
    c0, g0 = r.stats["cost_evals"], r.stats["grad_evals"]
    t0 = time.perf_counter()
    if pr:
        pr.enable()
    for _ in range(steps):
        r.step(w["N"], w["M"], save_update=True)
    if pr:
        pr.disable()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"{name}: {dt * 1e3:.3f} ms per Robot.step()  ({(r.stats['cost_evals'] - c0) / steps:.1f} cost evals, "
          f"{(r.stats['grad_evals'] - g0) / steps:.1f} gradient evals per step)")
    if pr:
        pstats.Stats(pr).sort_stats("cumulative").print_stats(35)


if __name__ == "__main__":
    for name in sys.argv[1:] or ["c1", "c2"]:
        run(name)
        run(name, profile=True)

#!/usr/bin/env python
"""Where a Robot.step() goes at the small configurations: CUDA events and host clocks at the phase boundaries of
kldiv_planner (no extra synchronisation: the events ride in the stream).  Usage: step_phases.py c2 [c1]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
import workloads as wl  # noqa: E402
from control_torch import engine  # noqa: E402
from control_torch.klerg import Robot  # noqa: E402

marks = []


class range_probe:
    """Stands in for engine.nvtx_range: an event + a host time stamp at the start of every named phase."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((self.name, time.perf_counter(), ev))

    def __exit__(self, *a):
        return False


def run(name, steps=40):
    steps = 6 if wl.WORKLOADS[name]["N"] >= 1_000_000 else steps
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
    engine.nvtx_range = range_probe
    import control_torch.klerg as kk
    kk.engine.nvtx_range = range_probe
    acc_host, acc_gpu = {}, {}
    fn_host, fn_gpu = {}, {}

    def wrap(obj, attr, label, gpu=False):
        """Host time inside obj.attr per step; gpu=True: also the stream time between its first and last launch."""
        f = getattr(obj, attr)

        def g(*a, **k):
            if gpu:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            t0 = time.perf_counter()
            out = f(*a, **k)
            fn_host[label] = fn_host.get(label, 0.0) + time.perf_counter() - t0
            if gpu:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                fn_gpu.setdefault(label, []).append((e0, e1))
            return out
        setattr(obj, attr, g)

    from control_torch.planner import PlannerContext
    wrap(PlannerContext, "optimize", "ctx.optimize (enqueue the loop)", gpu=True)
    wrap(r, "save_update", "save_update")
    wrap(r.robot, "step", "robot.step (dynamics)")
    wrap(r, "_after_plan", "prefetch launch")
    wrap(kk.engine, "device_uniform", "device_uniform")
    wrap(r.memory_buffer, "sample_device", "buffer.sample_device")
    wrap(r, "_pdf", "target pdf", gpu=True)
    wrap(kk.engine, "target_weight", "target_weight", gpu=True)
    wrap(kk.engine, "footprint_sum_max", "history sum+max", gpu=True)
    wrap(r, "_optimize_on_device", "_optimize_on_device (enqueue + read-back)")
    for _ in range(steps):
        marks.clear()
        with range_probe("step.begin"):
            pass
        r.step(w["N"], w["M"], save_update=True)
        with range_probe("step.end"):
            pass
        torch.cuda.synchronize()
        for (n0, t0, e0), (n1, t1, e1) in zip(marks[:-1], marks[1:]):
            acc_host[n0] = acc_host.get(n0, 0.0) + (t1 - t0)
            acc_gpu[n0] = acc_gpu.get(n0, 0.0) + e0.elapsed_time(e1) * 1e-3
    print(f"{name}: phase (from its start to the next phase's start): host ms | stream ms, per step")
    for k in acc_host:
        print(f"  {k:28s} {acc_host[k] / steps * 1e3:7.3f} | {acc_gpu[k] / steps * 1e3:7.3f}")
    print(f"  {'total':28s} {sum(acc_host.values()) / steps * 1e3:7.3f} | {sum(acc_gpu.values()) / steps * 1e3:7.3f}")
    print("  inside: host ms per step | stream ms between the call's first and last launch")
    for k, v in fn_host.items():
        gpu = sum(a.elapsed_time(b) for a, b in fn_gpu.get(k, [])) / steps if k in fn_gpu else float("nan")
        print(f"    {k:44s} {v / steps * 1e3:7.3f} | {gpu:7.3f}")


if __name__ == "__main__":
    for name in sys.argv[1:] or ["c2", "c1"]:
        run(name)

#!/usr/bin/env python
"""History footprint + spread pass (klerg_footprint_sum_max) at a workload's full size: tensor-core form against the
CUDA-core form, CUDA events.  Usage: time_footprint.py [c4] [n_samples] [n_rows]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
import workloads as wl  # noqa: E402
from control_torch import _cabi as cabi, engine  # noqa: E402
from control_torch.klerg import Robot  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = wl.WORKLOADS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else w["N"]
m = int(sys.argv[3]) if len(sys.argv) > 3 else w["M"]
lims = [wl.LIMS[s] for s in w["states"]]
dev = torch.device("cuda")
target = wl.make_target("gmm", lims, seed=1, device=dev)
probe = Robot(**wl.robot_kwargs(name, target, n_samples=n))
D = len(w["states"])
spec = cabi.kernel_spec(D, probe.planner.num_states, probe.explr_locs.tolist(), probe.std.tolist(), 1.0)
g = torch.Generator(device=dev).manual_seed(0)
lo = (torch.tensor([a for a, _ in lims]) * 1.15).to(dev)
hi = (torch.tensor([b for _, b in lims]) * 1.15).to(dev)
smp = lo + torch.rand(n, D, generator=g, device=dev) * (hi - lo)
packed = engine.pack_samples(spec, smp)
hist = wl.random_walk_history(name, m).to(dev)
t_sum = int(float(os.environ.get("TSUM_FRAC", "1")) * m)
res = {}
for tc in (False, True):
    for _ in range(2):
        out = engine.footprint_sum_max(spec, hist, t_sum, packed, n, tensor_cores=tc)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 3
    for _ in range(reps):
        out = engine.footprint_sum_max(spec, hist, t_sum, packed, n, tensor_cores=tc)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    res[tc] = out
    flag = ""
    if tc:
        sc = engine._tc_scratch[(torch.cuda.current_device(), cabi.raw_stream())]
        flag = f" (radius flag {int(sc[32:36].view(torch.int32).item())})"
    print(f"{name}: {m} rows x {n} samples, {'tensor cores' if tc else 'CUDA cores  '}: {ms:8.2f} ms = {m * n / ms * 1e3:.3e} pairs/s{flag}")
s0, m0, _ = res[False]
s1, m1, _ = res[True]
rel = ((s1[:n] - s0[:n]).abs() / s0[:n].abs().clamp_min(1e-30)).max().item()
relm = ((m1[:n] - m0[:n]).abs() / m0[:n].abs().clamp_min(1e-30)).max().item()
print(f"max relative difference: sum {rel:.2e}, max {relm:.2e}")

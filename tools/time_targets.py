"""Timing of BASELINE config 5 (16 belief targets, 1e6 samples): one fused K-target launch vs 16 single evals."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200"), os.path.join(ROOT, "tests")]
import torch
import workloads as wl
from test_gpu_full_size import build

s = build("c5", 1_000_000, 3_000)
ctx = s["ctx_for"](s["samples"], s["p_raw"], 1_000_000)
ctx.set_history(s["hist"])
e = s["engine"]
P = torch.stack([wl.make_target("gmm", s["lims"], seed=20 + k, device=s["dev"]).pdf_torch(s["samples"]) for k in range(16)])
stats = torch.stack([e.vector_stats(P[k].contiguous())[:1] for k in range(16)])
ctx.set_targets(P.contiguous(), stats)
u = wl.random_controls((s["H"], s["D"]), seed=4).to(s["dev"])
for fused in (True, False):
    ctx.fused = fused
    for _ in range(3):
        ctx.gradient_targets(u)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ctx.gradient_targets(u)
    b.record()
    torch.cuda.synchronize()
    print("fused K-target launch" if fused else "16 separate unfused evals", a.elapsed_time(b) / 10, "ms per 16-target gradient")
ctx.fused = True
ctx.targets_path = "tensor"
for _ in range(3):
    ctx.gradient_targets(u)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ctx.gradient_targets(u)
b.record()
torch.cuda.synchronize()
print("tensor-core shared-psi path (rollout + forward + targets gradient + adjoints)", a.elapsed_time(b) / 10, "ms per 16-target gradient")
ctx.targets_path = "fused"
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    for k in range(16):
        ctx.set_target(ctx.P[k, :ctx.n].contiguous(), ctx.P_stats[k:k + 1])
        ctx.gradient(u)
b.record()
torch.cuda.synchronize()
print("16 fused single-target evals", a.elapsed_time(b) / 10, "ms")

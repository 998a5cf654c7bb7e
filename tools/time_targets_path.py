"""Per-launch timing of the tensor-core K-target path at config 5 (rollout, forward pass, contraction, adjoints)."""
import os, sys
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


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


ro = e.rollout(ctx.dyn, ctx.bar, ctx.x0, u, R0=ctx.R0, want_lin=True)
traj = ro["traj"][0]
pre = traj[: ctx.H]
v, totals = e.footprint(ctx.spec, 0, pre, ctx.packed, ctx.n, add_in=ctx.q_base)
tw = totals.unsqueeze(0)
gp, _ = e.kl_gradient_targets(ctx.spec, pre, ctx.packed, ctx.n, v[0], tw, ctx.P, ctx.floor)
print("rollout            %.1f us" % timed(lambda: e.rollout(ctx.dyn, ctx.bar, ctx.x0, u, R0=ctx.R0, want_lin=True)))
print("forward footprint  %.1f us" % timed(lambda: e.footprint(ctx.spec, 0, pre, ctx.packed, ctx.n, add_in=ctx.q_base)))
print("targets gradient   %.1f us" % timed(lambda: e.kl_gradient_targets(ctx.spec, pre, ctx.packed, ctx.n, v[0], tw, ctx.P, ctx.floor)))
print("adjoint x16        %.1f us" % timed(lambda: e.adjoint_targets(ctx.dyn, ctx.spec, gp.unsqueeze(1), ro["dbarr"][0], None, traj, u, ctx.rinv, ctx.alpha, ctx.ctrl_lo, ctx.ctrl_hi)))
ctx.targets_path = "tensor"
print("whole path         %.1f us" % timed(lambda: ctx.gradient_targets(u)))

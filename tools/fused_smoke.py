"""Debug helper: one fused cost + gradient eval on random inputs, printing library errors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
import torch
import workloads as wl
from control_torch import _cabi as cabi, engine
from control_torch.klerg import Robot
from control_torch.planner import PlannerContext

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else wl.WORKLOADS[name]["N"]
w = wl.WORKLOADS[name]
lims = [wl.LIMS[s] for s in w["states"]]
dev = torch.device("cuda")
target = wl.make_target("gmm", lims, seed=1, device=dev)
kw = wl.robot_kwargs(name, target, n_samples=n)
probe = Robot(**kw)
D, H = len(w["states"]), w["H"]
res = {}
for fused in (False, True):
    ctx = PlannerContext(probe.planner.spec, probe.barrier.spec(), probe.explr_locs.tolist(), H,
                         torch.diagonal(probe.R_inv).tolist(), probe.control_lim[:, 0].tolist(),
                         probe.control_lim[:, 1].tolist(), alpha=1.0, fused=fused)
    g = torch.Generator().manual_seed(0)
    lo = torch.tensor([a for a, _ in lims]) * 1.15
    hi = torch.tensor([b for _, b in lims]) * 1.15
    smp = (lo + torch.rand(n, D, generator=g) * (hi - lo)).to(dev)
    ctx.set_samples(smp, probe.std.tolist(), 1.0)
    ctx.set_state(torch.tensor(kw["x0"], dtype=torch.float32, device=dev))
    p_raw = target.pdf_torch(smp).contiguous()
    p, p_stats, _ = engine.target_weight(2, smp, lo.tolist(), hi.tolist(), None, p_raw, n, 1.0, True)
    ctx.set_target(p, p_stats)
    ctx.set_history(wl.random_walk_history(name, min(w["M"], 3000)).to(dev))
    U = wl.random_controls((5, H, D), seed=3).to(dev)
    c = ctx.costs(U)
    gr = ctx.gradient(U[0], keep=True)
    torch.cuda.synchronize()
    if fused and os.environ.get("PROFILE"):
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        ctx.gradient(U[0])
        ctx.costs(U)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    if fused:
        for _ in range(5):
            ctx.gradient(U[0])
        torch.cuda.synchronize()
        st = engine.debug_stamps()
        names = ["rollout(states)", "forward", "meet1", "gradient", "row sums", "lin+gather", "adjoint+end"]
        print("finisher CTA, phase cycles:", {k: st[i + 1] - st[i] for i, k in enumerate(names)}, "total", st[7],
              "| inside rollout phase: stage u/x0", st[8], "rollout_block", st[9] - st[8], "x2 staging", st[1] - st[9])
    res[fused] = dict(cost=c.cpu(), du=gr["du"].cpu(), dj=gr["djdlam"].cpu(), us=gr["u_star"].cpu(), dgdx=gr["dgdx"].cpu(),
                      v=gr["v"][:n].cpu(), tot=gr["totals"].cpu())
for k in res[True]:
    a, b = res[True][k].double(), res[False][k].double()
    print(k, "max rel err", float(((a - b).abs() / (b.abs().max() + 1e-30)).max()))
print("cost", res[True]["cost"].tolist())

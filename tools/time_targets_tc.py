"""Time the shared-psi tensor-core gradient (klerg_kl_gradient_targets) at the config-5 size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
import torch
import workloads as wl
from control_torch import _cabi as cabi, engine

states, H, K, n = "xyz", 50, 16, int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
D = len(states)
g = torch.Generator().manual_seed(1)
lims = torch.tensor([wl.LIMS[c] for c in states])
lo, hi = lims[:, 0] * 1.15, lims[:, 1] * 1.15
samples = (lo + torch.rand(n, D, generator=g) * (hi - lo)).cuda()
std = wl.std_from_ratio([wl.LIMS[c] for c in states], n)
spec = cabi.kernel_spec(D, 2 * D, list(range(D)), [std] * D, 1.0)
packed = engine.pack_samples(spec, samples)
walk = torch.cumsum(torch.randn(H, D, generator=g) * 0.06, 0).clamp(-0.9, 0.9)
traj = torch.hstack([walk, torch.zeros(H, D)]).float().cuda().contiguous()
q_base = (torch.rand(n, generator=g) * 0.5).cuda()
v, totals = engine.footprint(spec, 0, traj, packed, n, add_in=q_base)
totals_w = totals.unsqueeze(0)
P = torch.zeros((K, packed.shape[1]), device="cuda")
for k in range(K):
    P[k, :n] = wl.make_target("gmm", [wl.LIMS[c] for c in states], seed=40 + k, device="cuda").pdf_torch(samples)

def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

t_tc = timed(lambda: engine.kl_gradient_targets(spec, traj, packed, n, v[0], totals_w, P))
p0 = P[0, :n].contiguous()
t_one = timed(lambda: engine.kl_gradient_fused(spec, traj, packed, n, v[0], totals_w, p0))
print(f"N={n} H={H} K={K}: shared-psi tensor-core gradient {t_tc*1e3:.1f} us for {K} targets "
      f"({K*H*n/t_tc/1e9:.2f}e12 target-pairs/s); legacy per-target FP32 kernel {t_one*1e3:.1f} us per target; fault {engine.targets_gradient_fault()}")

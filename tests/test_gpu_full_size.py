"""BASELINE.json's FULL sizes on the GPU, checked through size-independent properties (the CPU oracle
would need hours there): additivity of the footprint over history splits, exactness of the spread
maximum over splits, sample-split additivity of totals and gradient, candidate-order invariance of
the batched costs, linearity of the gradient in the target density."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import workloads as wl  # noqa: E402


def build(name, n, m, H=None, seed=0):
    from control_torch import _cabi as cabi, engine
    from control_torch.klerg import Robot
    from control_torch.planner import PlannerContext
    w = wl.WORKLOADS[name]
    H = H or w["H"]
    lims = [wl.LIMS[s] for s in w["states"]]
    D = len(lims)
    dev = torch.device("cuda")
    target = wl.make_target("gmm", lims, seed=1, device=dev)
    kw = wl.robot_kwargs(name, target, n_samples=n, horizon=H, cap=max(m, 8))
    torch.manual_seed(seed)
    probe = Robot(**kw)
    lo = (torch.tensor([a for a, _ in lims]) * 1.15).to(dev)
    hi = (torch.tensor([b for _, b in lims]) * 1.15).to(dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    samples = lo + torch.rand(n, D, generator=g, device=dev) * (hi - lo)
    hist = wl.random_walk_history(name, m, seed=seed).to(dev)

    def ctx_for(smp, p_raw, n_total):
        ctx = PlannerContext(probe.planner.spec, probe.barrier.spec(), probe.explr_locs.tolist(), H,
                             torch.diagonal(probe.R_inv).tolist(), probe.control_lim[:, 0].tolist(),
                             probe.control_lim[:, 1].tolist(), alpha=1.0)
        ctx.set_samples(smp, probe.std.tolist(), 1.0)
        ctx.set_state(torch.tensor(kw["x0"], dtype=torch.float32, device=dev))
        p, p_stats, _ = engine.target_weight(2, smp, lo.tolist(), hi.tolist(), None, p_raw, n_total, 1.0, True)
        ctx.set_target(p, p_stats)
        return ctx

    p_raw = torch.cat([target.pdf_torch(c) for c in samples.split(1_000_000)]).contiguous()
    return dict(ctx_for=ctx_for, samples=samples, p_raw=p_raw, hist=hist, probe=probe, D=D, H=H, dev=dev, engine=engine,
                cabi=cabi, target=target, lims=lims)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def test_config4_full_history_and_spread_split_properties():
    """1e7 samples x 1e5 history states, 6-D: sum over a split history adds up, max over a split is exact."""
    s = build("c4", 10_000_000, 100_000)
    ctx = s["ctx_for"](s["samples"], s["p_raw"], 10_000_000)
    e, spec, n = s["engine"], ctx.spec, ctx.n
    whole, tot = e.footprint(spec, 0, s["hist"], ctx.packed, n)
    a, tot_a = e.footprint(spec, 0, s["hist"][:37_003], ctx.packed, n)
    b, tot_b = e.footprint(spec, 0, s["hist"][37_003:], ctx.packed, n)
    assert rel(a[0, :n] + b[0, :n], whole[0, :n]) < 1e-5  # fp32 sums of up to 1e5 terms, two association orders
    assert abs(float(tot_a[0, 0] + tot_b[0, 0]) / float(tot[0, 0]) - 1) < 1e-6
    mw, _ = e.footprint(spec, 1, s["hist"], ctx.packed, n)
    ma, _ = e.footprint(spec, 1, s["hist"][:37_003], ctx.packed, n)
    mb, _ = e.footprint(spec, 1, s["hist"][37_003:], ctx.packed, n)
    # a maximum does not round, but every staged chunk of rows is evaluated around its own centre and the chunk
    # boundaries of the split lists differ from the whole list's
    assert rel(torch.maximum(ma[0, :n], mb[0, :n]), mw[0, :n]) < 5e-5
    assert float(mw[0, :n].max()) <= 1.0 and float(mw[0, :n].min()) >= 0.0


def test_config4_full_eval_is_additive_over_sample_halves():
    """The fused eval of the 1e7-sample workspace: sum of q and the gradient are sums over samples, so the two
    halves of the workspace (each with the WHOLE workspace's normalisers) add up to the whole."""
    s = build("c4", 10_000_000, 2_000)
    n, half = 10_000_000, 5_000_000
    ctx = s["ctx_for"](s["samples"], s["p_raw"], n)
    ctx.set_history(s["hist"])
    u = wl.random_controls((s["H"], s["D"]), seed=3).to(s["dev"])
    g = ctx.gradient(u, keep=True)
    tot = g["totals"].reshape(-1).clone()
    dgdx = g["dgdx"].clone()
    # explicit-weight gradient on each half with the whole-workspace importance ratio
    e, spec = s["engine"], ctx.spec
    q = ctx.q_from(g["v"], g["totals"])
    w = (ctx.p[:n] / q).contiguous()
    pre = g["traj"].contiguous()
    parts = []
    for lo_i, hi_i in ((0, half), (half, n)):
        smp = s["samples"][lo_i:hi_i].contiguous()
        packed = e.pack_samples(spec, smp)
        parts.append(e.kl_gradient(spec, pre, packed, hi_i - lo_i, w[lo_i:hi_i].contiguous()))
        v_half, t_half = e.footprint(spec, 0, pre, packed, hi_i - lo_i, add_in=ctx.q_base[lo_i:hi_i].contiguous())
        parts.append(t_half[0, 0])
    assert rel(parts[0] + parts[2], dgdx) < 2e-4  # two independent fp32 summation orders over 1e7 samples
    assert abs(float(parts[1] + parts[3]) / float(tot[0]) - 1) < 1e-6
    assert not s["engine"].fused_fault()


def test_config3_full_batch_is_order_invariant():
    """1024 candidates x H=50 x 1e6 samples: the batch equals any reordering of itself, bit for bit."""
    s = build("c3", 1_000_000, 3_000)
    ctx = s["ctx_for"](s["samples"], s["p_raw"], 1_000_000)
    ctx.set_history(s["hist"])
    u0 = wl.random_controls((s["H"], s["D"]), seed=1)
    g = torch.Generator().manual_seed(2)
    U = (u0.unsqueeze(0) + 0.1 * torch.randn(1024, s["H"], s["D"], generator=g)).to(s["dev"])
    c = ctx.costs(U)
    perm = torch.randperm(1024, generator=g).to(s["dev"])
    assert torch.equal(ctx.costs(U[perm].contiguous()), c[perm])
    assert torch.isfinite(c).all() and float(c.std()) > 0


@pytest.mark.parametrize("path", ["fused", "tensor"])
def test_config5_full_gradient_is_linear_in_the_target(path):
    """16 belief targets over 1e6 samples: q does not depend on p, so dgdx(p_j + p_k) = dgdx(p_j) + dgdx(p_k); and the
    two implementations (pair pass per target / shared-psi tensor-core contraction) agree."""
    s = build("c5", 1_000_000, 3_000)
    ctx = s["ctx_for"](s["samples"], s["p_raw"], 1_000_000)
    ctx.targets_path = path
    ctx.set_history(s["hist"])
    e = s["engine"]
    P = torch.stack([wl.make_target("gmm", s["lims"], seed=20 + k, device=s["dev"]).pdf_torch(s["samples"]) for k in range(16)])
    P = torch.cat([P, (P[3] + P[11]).unsqueeze(0)])
    stats = torch.stack([e.vector_stats(P[k].contiguous())[:1] for k in range(17)])
    ctx.set_targets(P.contiguous(), stats)
    u = wl.random_controls((s["H"], s["D"]), seed=4).to(s["dev"])
    g = ctx.gradient_targets(u)
    assert g["dgdx"].shape == (17, s["H"], 2 * s["D"])
    assert rel(g["dgdx"][3] + g["dgdx"][11], g["dgdx"][16]) < 2e-5
    if path == "tensor":
        assert not e.targets_gradient_fault()
        ctx.targets_path = "fused"
        g2 = ctx.gradient_targets(u)
        for key in ("dgdx", "du", "djdlam"):
            assert rel(g[key], g2[key]) < 1e-4, key
    costs = ctx.costs_targets(u.unsqueeze(0))
    assert costs.shape == (17, 1) and torch.isfinite(costs).all()

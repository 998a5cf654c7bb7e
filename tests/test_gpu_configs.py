"""Scaled-down BASELINE configs 3, 4, 5 on the GPU against the CPU oracle (tolerance 1e-4 relative,
BASELINE.json north_star): batched candidate costs, 6-D pose workspace with a long history, K belief
targets.  Full sizes are covered by size-independent properties (candidate-order invariance, the
batch equals the one-by-one evaluation, sharding invariance in test_sharding.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import workloads as wl  # noqa: E402
from oracle import klerg_oracle as ko  # noqa: E402

RTOL = 1e-4


def close(a, b, rtol=RTOL, atol_frac=0.0, what=""):
    a = torch.as_tensor(a).detach().double().cpu().numpy()
    b = torch.as_tensor(b).detach().double().cpu().numpy()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_frac * (np.abs(b).max() if b.size else 0) + 1e-37, err_msg=what)


def setup(name, n, m, H=None, seed=0):
    from control_torch import engine
    from control_torch.klerg import Robot
    from control_torch.planner import PlannerContext
    w = wl.WORKLOADS[name]
    H = H or w["H"]
    lims = [wl.LIMS[s] for s in w["states"]]
    D = len(lims)
    dev = torch.device("cuda")
    target = wl.make_target("gmm", lims, seed=1, device="cpu")
    kw = wl.robot_kwargs(name, target, n_samples=n, horizon=H, cap=max(m, 8))
    torch.manual_seed(seed)
    probe = Robot(**kw)
    oracle = ko.OracleRobot(**kw)
    g = torch.Generator().manual_seed(seed)
    lo = torch.tensor([a for a, _ in lims]) * 1.15
    hi = torch.tensor([b for _, b in lims]) * 1.15
    samples = lo + torch.rand(n, D, generator=g) * (hi - lo)
    hist = wl.random_walk_history(name, m, seed=seed)
    p_raw = target.pdf_torch(samples)
    p = ko.renormalize(p_raw.clone())
    q_base = ko.footprint_sum(hist, samples, oracle.explr_locs, oracle.std, torch.ones(1))
    ctx = PlannerContext(probe.planner.spec, probe.barrier.spec(), probe.explr_locs.tolist(), H,
                         torch.diagonal(probe.R_inv).tolist(), probe.control_lim[:, 0].tolist(),
                         probe.control_lim[:, 1].tolist(), alpha=1.0)
    ctx.set_samples(samples.to(dev), probe.std.tolist(), 1.0)
    ctx.set_state(torch.tensor(kw["x0"], dtype=torch.float32, device=dev))
    ctx.set_target(p.to(dev), engine.vector_stats(p.to(dev))[:1].contiguous())
    ctx.set_history(hist.to(dev))
    close(ctx.q_base[:n], q_base, what="history footprint")
    return dict(ctx=ctx, oracle=oracle, samples=samples, p=p, q_base=q_base, D=D, H=H, dev=dev, engine=engine)


def oracle_grad(o, samples, p, q_base, u):
    o.u = u.clone()
    _, lin, traj = o.forward(0)
    q = ko.renormalize(q_base + ko.footprint_sum(traj, samples, o.explr_locs, o.std, torch.ones(1)))
    du, dj = o.backward(samples, p.clone(), q, torch.ones(1), lin, traj)
    return du, dj


def test_config3_batched_candidates():
    """64 candidates x H=50 x 2e4 samples in ONE launch (klerg_eval_costs_batch); 6 of them against the oracle."""
    s = setup("c3", 20_000, 300)
    u0 = wl.random_controls((s["H"], s["D"]), seed=1)
    g = torch.Generator().manual_seed(2)
    U = u0.unsqueeze(0) + 0.1 * torch.randn(64, s["H"], s["D"], generator=g)
    got = s["ctx"].costs(U.to(s["dev"])).cpu()
    for b in (0, 7, 8, 31, 62, 63):
        want = s["oracle"].get_cost(s["samples"], s["p"].clone(), s["q_base"], U[b])
        close(got[b], want.reshape(()), what=f"candidate {b}")
    # the batch is the one-by-one evaluation, in any order
    perm = torch.randperm(64, generator=g)
    got_perm = s["ctx"].costs(U[perm].to(s["dev"])).cpu()
    assert torch.equal(got_perm, got[perm])
    # one by one through the <= 8-candidate kernel: same numbers up to the rounding of the sums (other grid, other centre
    # of the expanded pair form)
    one = torch.stack([s["ctx"].costs(U[b:b + 1].to(s["dev"]))[0] for b in (3, 40)]).cpu()
    close(one, got[[3, 40]], rtol=2e-6, what="batch vs one by one")
    # and through launches of 8 (the path of a sharded context): same again
    s["ctx"].batch_costs = False
    eight = s["ctx"].costs(U.to(s["dev"])).cpu()
    s["ctx"].batch_costs = True
    close(eight, got, rtol=2e-6, what="batch vs launches of 8")


@pytest.mark.parametrize("name,n,B", [("c3", 1003, 9), ("c3", 2048, 17), ("c3", 2050, 12), ("c3", 40_001, 33),
                                      ("c1", 777, 10), ("c4", 9_999, 11)])
def test_batch_costs_ragged_sizes(name, n, B):
    """klerg_eval_costs_batch on sizes around its work-unit boundaries (fewer samples than one pass of the CTA, exactly
    one chunk, a two-sample tail, B not a multiple of 8; D = 2, 3, 6) against launches of 8 through eval_cost_kernel,
    and directly for B < 8."""
    s = setup(name, n, 200, H=12 if name == "c1" else 20)
    ctx, dev = s["ctx"], s["dev"]
    g = torch.Generator().manual_seed(B)
    U = (wl.random_controls((s["H"], s["D"]), seed=3).unsqueeze(0) + 0.1 * torch.randn(B, s["H"], s["D"], generator=g)).to(dev)
    got = ctx.costs(U).cpu()
    ctx.batch_costs = False
    want = ctx.costs(U).cpu()
    ctx.batch_costs = True
    close(got, want, rtol=5e-6, what="batch vs launches of 8")
    for b in (0, B - 1):
        ref = s["oracle"].get_cost(s["samples"], s["p"].clone(), s["q_base"], U[b].cpu())
        close(got[b], ref.reshape(()), rtol=5e-4 if name == "c4" else 1e-4, what=f"candidate {b} vs oracle")
    few, fault = s["engine"].eval_costs_batch(ctx.spec, ctx.dyn, ctx.bar, ctx.x0, ctx.R0, U[:3].contiguous(), ctx.packed, ctx.n,
                                              ctx.q_base, ctx.p, ctx.p_stats, ctx.floor)
    close(few.cpu(), want[:3], rtol=5e-6, what="3 candidates through the batch kernel")
    assert float(fault[0]) == 0.0


@pytest.mark.parametrize("H", [20, 48, 50, 64])
def test_config4_pose_workspace_long_history(H):
    """6-D pose (roll dynamics), 3e4 samples, 2000 history states: cost + gradient eval vs the oracle.
    H=50 exercises the mixed schedule (14 warps own 3 states, 2 warps own 4), H=20 the uniform one, H=48 / 64 the
    all-wide schedules of horizons that are multiples of 16 (3 / 4 states per warp)."""
    s = setup("c4", 30_000, 2_000, H=H)
    U = wl.random_controls((3, s["H"], s["D"]), seed=5)
    got = s["ctx"].costs(U.to(s["dev"])).cpu()
    ctx = s["ctx"]
    bar_gpu = s["engine"].rollout(ctx.dyn, ctx.bar, ctx.x0, U.to(s["dev"]), R0=ctx.R0)["barrier"].cpu()
    for b in range(3):
        s["oracle"].trace = []
        want = s["oracle"].get_cost(s["samples"], s["p"].clone(), s["q_base"], U[b])
        dkl = float(s["oracle"].trace[-1]["dkl"])
        close(got[b], want.reshape(()), rtol=5e-4, what=f"cost {b}")  # quartic barrier amplifies the fp32 matrix_exp difference
        # ... and only the barrier: the KL term alone (cost minus the wall barrier of the same rollout) holds the 1e-4 of
        # the north_star; the absolute term covers the fp32 subtraction of the two parts
        np.testing.assert_allclose(float(got[b]) - float(bar_gpu[b]), dkl, rtol=1e-4, atol=5e-7 * abs(float(want)),
                                   err_msg=f"KL part of cost {b}")
        np.testing.assert_allclose(float(bar_gpu[b]), float(want) - dkl, rtol=2e-3, atol=5e-7 * abs(float(want)),
                                   err_msg=f"barrier part of cost {b}")
    s["oracle"].trace = None
    g = s["ctx"].gradient(U[0].to(s["dev"]))
    du, dj = oracle_grad(s["oracle"], s["samples"], s["p"], s["q_base"], U[0])
    close(g["du"], du, rtol=RTOL, atol_frac=5e-5, what="du")
    close(g["djdlam"], dj, rtol=RTOL, atol_frac=5e-5, what="djdlam")


@pytest.mark.parametrize("path", ["fused", "tensor"])
def test_config5_belief_targets(path):
    """4 belief targets over one workspace: per-target cost and gradient vs the oracle, target by target - through the
    fused launch (pair pass per target) and through the shared-psi tensor-core contraction."""
    s = setup("c5", 20_000, 300, H=20)
    s["ctx"].targets_path = path
    lims = [wl.LIMS[c] for c in wl.WORKLOADS["c5"]["states"]]
    P = torch.stack([ko.renormalize(wl.make_target("gmm", lims, seed=10 + k).pdf_torch(s["samples"])) for k in range(4)])
    Pd = P.to(s["dev"])
    stats = torch.stack([s["engine"].vector_stats(Pd[k])[:1] for k in range(4)])
    s["ctx"].set_targets(Pd, stats)
    U = wl.random_controls((2, s["H"], s["D"]), seed=9)
    costs = s["ctx"].costs_targets(U.to(s["dev"])).cpu()
    grads = s["ctx"].gradient_targets(U[0].to(s["dev"]))
    assert costs.shape == (4, 2) and grads["du"].shape == (4, s["H"], s["D"])
    for k in range(4):
        for b in range(2):
            want = s["oracle"].get_cost(s["samples"], P[k].clone(), s["q_base"], U[b])
            close(costs[k, b], want.reshape(()), what=f"target {k} cost {b}")
        du, dj = oracle_grad(s["oracle"], s["samples"], P[k], s["q_base"], U[0])
        close(grads["du"][k], du, rtol=RTOL, atol_frac=5e-5, what=f"target {k} du")
        close(grads["djdlam"][k], dj, rtol=RTOL, atol_frac=5e-5, what=f"target {k} djdlam")


@pytest.mark.parametrize("states,H,K,n", [("xyz", 50, 16, 40_000), ("xy", 20, 3, 5_003), ("xyzrpw", 33, 18, 9_000), ("xyz", 64, 32, 700)])
def test_shared_psi_targets_gradient_vs_per_target_kernel(states, H, K, n):
    """klerg_kl_gradient_targets (psi once per pair, tensor-core contraction over the samples) against the per-target
    FP32 kernel klerg_kl_gradient_fused on the same inputs: gradient partials and KL terms of every target."""
    from control_torch import _cabi as cabi
    from control_torch import engine
    D = len(states)
    g = torch.Generator().manual_seed(H * K + n)
    lims = torch.tensor([wl.LIMS[c] for c in states])
    lo, hi = lims[:, 0] * 1.15, lims[:, 1] * 1.15
    samples = (lo + torch.rand(n, D, generator=g) * (hi - lo)).cuda()
    std = wl.std_from_ratio([wl.LIMS[c] for c in states], n)
    spec = cabi.kernel_spec(D, 2 * D, list(range(D)), [std] * D, 1.0)
    packed = engine.pack_samples(spec, samples)
    # a wandering trajectory inside the workspace, zero velocities
    mid, half = lims.mean(1), (lims[:, 1] - lims[:, 0]) / 2
    walk = torch.cumsum(torch.randn(H, D, generator=g) * 0.06, 0).clamp(-0.9, 0.9)
    traj = torch.hstack([mid + half * walk, torch.zeros(H, D)]).float().cuda().contiguous()
    q_base = (torch.rand(n, generator=g) * 0.5).cuda()
    v, totals = engine.footprint(spec, 0, traj, packed, n, add_in=q_base)
    totals_w = totals.unsqueeze(0) if totals.dim() == 2 else totals
    P = torch.zeros((K, packed.shape[1]), device="cuda")
    for k in range(K):
        P[k, :n] = wl.make_target("gmm", [wl.LIMS[c] for c in states], seed=40 + k, device="cuda").pdf_torch(samples)
    gp, kl = engine.kl_gradient_targets(spec, traj, packed, n, v[0], totals_w, P)
    assert not engine.targets_gradient_fault()
    for k in range(K):
        gref, klref = engine.kl_gradient_fused(spec, traj, packed, n, v[0], totals_w, P[k, :n].contiguous())
        close(gp[k], gref, rtol=RTOL, atol_frac=5e-5, what=f"target {k} gradient partial")
        close(kl[k], klref, rtol=RTOL, what=f"target {k} KL terms")

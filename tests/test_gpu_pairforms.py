"""The two forms of the squared distance inside the pair passes, and the one-pass history footprint + spread.

Every pair pass evaluates |x - s|^2 either from the coordinate differences (exact form) or, where the staged state
set is narrow enough, in the expanded form |sc|^2 + |xc|^2 - 2 xc.sc around a centre of the set (fewer FP32 ops per
pair).  Checked here against the CPU oracle (1e-4 relative, BASELINE.json north_star): narrow sets (expanded form),
wide sets (kernel widths << extent: automatic fall-back to the exact form), the forced exact form
(KLERG_OPT_EXACT_PAIRS), and the two forms against each other."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import workloads as wl  # noqa: E402
from oracle import klerg_oracle as ko  # noqa: E402


def close(a, b, rtol=1e-4, atol_frac=0.0, what=""):
    a = torch.as_tensor(a).detach().double().cpu().numpy()
    b = torch.as_tensor(b).detach().double().cpu().numpy()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_frac * (np.abs(b).max() if b.size else 0) + 1e-37, err_msg=what)


@pytest.fixture
def exact_switch():
    from control_torch import _cabi as cabi
    lib = cabi.load()

    def set_exact(on):
        cabi.check(lib.klerg_set_option(cabi.OPT_EXACT_PAIRS, int(on)), "klerg_set_option")

    yield set_exact
    set_exact(False)


def _inputs(D, n, T, std, seed=0, walk=0.03):
    g = torch.Generator().manual_seed(seed)
    samples = torch.rand(n, D, generator=g) * 2.3 - 1.15
    pos = torch.cumsum(torch.randn(T, D, generator=g) * walk, 0)
    pos = torch.abs(((pos + 1) % 4) - 2) - 1  # reflected random walk inside [-1, 1]
    states = torch.hstack([pos, torch.zeros(T, D)]).contiguous()
    return samples, states, torch.full((D,), std)


@pytest.mark.parametrize("D,n,T,t_sum,std", [
    (3, 10_007, 2500, 900, 0.02), (6, 4_099, 1700, 1700, 0.08), (2, 3_001, 1030, 0, 0.01), (3, 2_050, 5, 3, 0.02),
    (3, 5_003, 2100, 1024, 2e-4),   # kernel width << path extent: every chunk falls back to the exact form
])
def test_sum_max_pass(D, n, T, t_sum, std, exact_switch):
    """klerg_footprint_sum_max = traj_footprint_vec over the first t_sum rows + traj_spread_vec over all rows."""
    from control_torch import _cabi as cabi, engine
    samples, states, scale = _inputs(D, n, T, std)
    explr = list(range(D))
    spec = cabi.kernel_spec(D, 2 * D, explr, scale.tolist(), 1.0)
    s_dev, st_dev = samples.cuda(), states.cuda()
    packed = engine.pack_samples(spec, s_dev)
    q, m, tot = engine.footprint_sum_max(spec, st_dev, t_sum, packed, n)
    want_q = ko.footprint_sum(states[:t_sum], samples, torch.tensor(explr), scale, torch.ones(1)) if t_sum else torch.zeros(n)
    want_m = ko.spread_max(states, samples, torch.tensor(explr), scale, torch.ones(1))
    # psi underflows for most pairs at the tiny widths: compare relative to the largest entry there
    frac = 1e-6 if std < 1e-3 else 1e-9
    close(q[:n], want_q, atol_frac=frac, what="sum over the drawn rows")
    close(m[:n], want_m, atol_frac=frac, what="max over all rows")
    close(tot[0], want_q.double().sum(), rtol=1e-5, what="total")
    # the separate passes give the same numbers up to the rounding of the expanded form (the chunks of staged rows,
    # hence their centres, differ where t_sum is not a multiple of the chunk size)
    q2, _ = engine.footprint(spec, 0, st_dev[:t_sum], packed, n)
    m2, _ = engine.footprint(spec, 1, st_dev, packed, n)
    close(q[:n], q2[0, :n], rtol=5e-5, atol_frac=1e-7, what="one pass vs separate sum pass")
    close(m[:n], m2[0, :n], rtol=5e-5, atol_frac=1e-7, what="one pass vs separate max pass")
    # forced exact form
    exact_switch(True)
    q3, m3, _ = engine.footprint_sum_max(spec, st_dev, t_sum, packed, n)
    exact_switch(False)
    close(q3[:n], want_q, atol_frac=frac, what="exact form: sum")
    close(m3[:n], want_m, atol_frac=frac, what="exact form: max")
    close(q[:n], q3[:n], rtol=1e-4, atol_frac=1e-6, what="expanded vs exact form")


def _eval_setup(name, n, m, H, std_scale=1.0, seed=0):
    from control_torch import engine
    from control_torch.klerg import Robot
    from control_torch.planner import PlannerContext
    w = wl.WORKLOADS[name]
    lims = [wl.LIMS[s] for s in w["states"]]
    D = len(lims)
    dev = torch.device("cuda")
    target = wl.make_target("gmm", lims, seed=1, device="cpu")
    kw = wl.robot_kwargs(name, target, n_samples=n, horizon=H, cap=max(m, 8))
    kw["std"] = kw["std"] * std_scale
    torch.manual_seed(seed)
    probe = Robot(**kw)
    oracle = ko.OracleRobot(**kw)
    g = torch.Generator().manual_seed(seed)
    lo = torch.tensor([a for a, _ in lims]) * 1.15
    hi = torch.tensor([b for _, b in lims]) * 1.15
    samples = lo + torch.rand(n, D, generator=g) * (hi - lo)
    hist = wl.random_walk_history(name, m, seed=seed)
    p = ko.renormalize(target.pdf_torch(samples).clone())
    q_base = ko.footprint_sum(hist, samples, oracle.explr_locs, oracle.std, torch.ones(1))
    ctx = PlannerContext(probe.planner.spec, probe.barrier.spec(), probe.explr_locs.tolist(), H,
                         torch.diagonal(probe.R_inv).tolist(), probe.control_lim[:, 0].tolist(),
                         probe.control_lim[:, 1].tolist(), alpha=1.0)
    ctx.set_samples(samples.to(dev), probe.std.tolist(), 1.0)
    ctx.set_state(torch.tensor(kw["x0"], dtype=torch.float32, device=dev))
    ctx.set_target(p.to(dev), engine.vector_stats(p.to(dev))[:1].contiguous())
    ctx.set_history(hist.to(dev))
    return dict(ctx=ctx, oracle=oracle, samples=samples, p=p, q_base=q_base, D=D, H=H, dev=dev, engine=engine)


@pytest.mark.parametrize("name,n,H,std_scale", [("c2", 30_011, 50, 1.0), ("c4", 20_003, 50, 1.0), ("c1", 3_000, 20, 1.0),
                                                ("c2", 30_011, 50, 0.002), ("c1", 3_000, 20, 0.001)])
def test_fused_evals_both_forms_vs_oracle(name, n, H, std_scale, exact_switch):
    """Cost and gradient evals against the oracle in the default mode (expanded form where the trajectory is narrow;
    the std_scale << 1 cases make it wide: automatic exact form) and with the exact form forced."""
    s = _eval_setup(name, n, 300, H, std_scale)
    ctx, o = s["ctx"], s["oracle"]
    U = 1.6 * wl.random_controls((3, H, s["D"]), seed=11)
    for forced in (False, True):
        exact_switch(forced)
        costs = ctx.costs(U.to(s["dev"])).cpu()
        for b in range(3):
            want = o.get_cost(s["samples"], s["p"].clone(), s["q_base"], U[b])
            # c4: the reference's fp32 matrix_exp differs from the device's Rodrigues form by ~3e-5 in the angles, which the
            # quartic wall amplifies (same tolerance as tests/test_gpu_configs.py)
            close(costs[b], want.reshape(()), rtol=5e-4 if name == "c4" else 2e-4, what=f"cost {b} forced_exact={forced}")
        g = ctx.gradient(U[0].to(s["dev"]), keep=True)
        o.u = U[0].clone()
        _, lin, traj = o.forward(0)
        q = ko.renormalize(s["q_base"] + ko.footprint_sum(traj, s["samples"], o.explr_locs, o.std, torch.ones(1)))
        du, dj = o.backward(s["samples"], s["p"].clone(), q, torch.ones(1), lin, traj)
        close(g["du"], du, rtol=1e-4, atol_frac=2e-5, what=f"du forced_exact={forced}")
        close(g["djdlam"], dj, rtol=1e-4, atol_frac=2e-5, what=f"djdlam forced_exact={forced}")
    exact_switch(False)
    assert not s["engine"].fused_fault()

"""Oracle vs the LIVE reference (only where /root/reference exists, i.e. in the authoring
container; skipped on the GPU box).  Complements test_oracle_golden.py, which checks the oracle
against the committed vectors the same reference produced."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference/franka_test/scripts"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")

from cases import ROBOT_CASES, MixtureTarget, robot_kwargs, seed_buffer_states  # noqa: E402
from oracle import klerg_oracle as ko  # noqa: E402


def _purge():
    for k in [k for k in sys.modules if k.split(".")[0] in ("control_torch", "franka")]:
        del sys.modules[k]


@pytest.fixture(scope="module")
def ref():
    """Import the reference's control_torch.  This repo's host mirror has the same package name (it is a
    drop-in), so it is taken off sys.path / sys.modules for the duration of the import and restored after."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import import_reference
    saved = list(sys.path)
    sys.path[:] = [p for p in sys.path if not p.rstrip("/").endswith("embodied-active-learning-vision_b200")]
    _purge()
    try:
        yield import_reference()  # the reference imports .dynamics lazily: keep it resolvable while the tests run
    finally:
        _purge()
        sys.path[:] = saved


@pytest.mark.parametrize("D,S,T,N", [(2, 4, 41, 333), (3, 6, 17, 1000), (6, 12, 9, 257)])
def test_pairwise_functions_match_live_reference(ref, D, S, T, N):
    _, ru, _, _, _ = ref
    g = torch.Generator().manual_seed(100 + D)
    traj = torch.rand(T, S, generator=g) * 2 - 1
    samples = torch.rand(N, D, generator=g) * 2.3 - 1.15
    explr = torch.arange(D)
    std = torch.rand(D, generator=g) * 0.1 + 0.03
    nu = torch.ones(1)
    w = torch.rand(N, generator=g) + 0.1
    assert torch.equal(ko.footprint_sum(traj, samples, explr, std, nu), ru.traj_footprint_vec(traj, samples, explr, std, nu))
    assert torch.equal(ko.spread_max(traj, samples, explr, std, nu), ru.traj_spread_vec(traj, samples, explr, std, nu))
    for x in traj[:3]:
        assert torch.equal(ko.kl_gradient(x, samples, explr, std, w, nu), ru.kldiv_grad_vec(x, samples, explr, std, w, nu))
    q = ru.traj_footprint_vec(traj, samples, explr, std, nu)
    np.testing.assert_allclose(ko.renormalize(q.clone()).numpy(), ru.renormalize(q.clone()).numpy(), rtol=2e-6)
    np.testing.assert_allclose(ko.unit_mass(q.clone()).numpy(), ru.cost_norm(q.clone()).numpy(), rtol=1e-6)


@pytest.mark.parametrize("name", ["xyz_small", "xyzrpw"])
def test_robot_steps_match_live_reference(ref, name):
    rk = ref[0]
    case = ROBOT_CASES[name]
    outs = []
    for cls in (rk.Robot, ko.OracleRobot):
        torch.manual_seed(4321)
        torch.set_num_threads(1)
        target = MixtureTarget(case["D"], seed=9)
        if case["states"] == "xyzrpw":
            target.mu[:, 3] = target.mu[:, 3] * 0.5 + 3.1
        r = cls(**robot_kwargs(case, target))
        r.test(case["n"])
        for s in seed_buffer_states(r.robot.state, case):
            r.memory_buffer.push(s)
        res = [r.step(case["n"], case["m"], save_update=True) for _ in range(4)]
        outs.append((res, r.u.clone(), r.last_plan.clone()))
    (ra, ua, pa), (rb, ub, pb) = outs
    for (sa, va, ca), (sb, vb, cb) in zip(ra, rb):
        np.testing.assert_allclose(sb, sa, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(cb, ca, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ub.numpy(), ua.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(pb.numpy(), pa.numpy(), rtol=1e-5, atol=1e-6)


def test_prior_dist_matches_live_reference(ref):
    rk = ref[0]
    assert "/root/reference" in rk.__file__
    x = torch.rand(300, 3, generator=torch.Generator().manual_seed(0)) * 3 - 1.5
    np.testing.assert_allclose(ko.OraclePrior("xyz").pdf_torch(x).numpy(), rk.PriorDist("xyz").pdf_torch(x).numpy(), rtol=1e-6)


def test_feedback_policies_match_live_reference(ref):
    """BarrierPush / LQR (default_policies.py:53-119): the reference classes, the oracle's rule and the host methods of
    the B200 mirror (loaded by file path: the package of that name on sys.path is the reference's here)."""
    import importlib.util
    rd = ref[3]
    import control_torch.default_policies as ref_pol
    assert "/root/reference" in ref_pol.__file__
    spec = importlib.util.spec_from_file_location("mirror_default_policies", os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "embodied-active-learning-vision_b200", "control_torch",
        "default_policies.py"))
    mp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mp)
    model = rd.DoubleIntegratorEnv(dt=0.2, x0=torch.zeros(6), states="xyz")
    a, b, o = ref_pol.BarrierPush(model, 10), mp.BarrierPush(model, 10), ko.OracleFeedback("BarrierPush", model, 10)
    g = torch.Generator().manual_seed(0)
    for k in range(200):
        x = torch.rand(6, generator=g) * 3 - 1.5
        if k % 5 == 0:
            x[0] = 1.0
        if k % 7 == 0:
            x[1] = -1.0
        u = torch.rand(12, 3, generator=g)
        for idx in (0, 2):
            a.reset(x, u.clone(), idx)
            b.reset(x, u.clone(), idx)
            ra, rb = a(x), b(x)
            want, dmu = o.act(x, u[0].clone() if o.uses_plan(idx) else torch.zeros(3))
            assert torch.equal(ra, rb) and torch.equal(ra, want)
            assert torch.equal(a.dx(x, ra), b.dx(x, rb)) and torch.equal(a.dx(x, ra), dmu)
    k_ref = ref_pol.LQR(model, 10).Klqr
    assert torch.allclose(k_ref, mp.LQR(model, 10).Klqr) and torch.allclose(k_ref, ko.OracleFeedback("LQR", model, 10).K)


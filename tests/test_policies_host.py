"""Host-side rule of the state-feedback default policies (BarrierPush, LQR: reference default_policies.py:53-119) in
the B200 mirror against the oracle's restatement; test_oracle_live.py checks both against the live reference classes."""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
from oracle import klerg_oracle as ko  # noqa: E402


def _mirror():
    spec = importlib.util.spec_from_file_location(
        "mirror_default_policies", os.path.join(ROOT, "embodied-active-learning-vision_b200", "control_torch", "default_policies.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _states(n=200):
    g = torch.Generator().manual_seed(0)
    for k in range(n):
        x = torch.rand(6, generator=g) * 3 - 1.5
        if k % 5 == 0:
            x[0] = 1.0   # exactly on the wall
        if k % 7 == 0:
            x[1] = -1.0
        yield x, torch.rand(12, 3, generator=g)


def test_mirror_policies_match_oracle_rule():
    mp = _mirror()
    model = ko.OracleDynamics("double", 0.2, torch.zeros(6).numpy(), "xyz", torch.float32)
    push, lqr = mp.BarrierPush(model, 10), mp.LQR(model, 10)
    o_push, o_lqr = ko.OracleFeedback("BarrierPush", model, 10), ko.OracleFeedback("LQR", model, 10)
    assert torch.allclose(lqr.Klqr, o_lqr.K)
    for x, u in _states():
        for idx in (0, 3):
            push.reset(x, u.clone(), idx)
            got = push(x)
            want, dmu = o_push.act(x, u[0].clone() if o_push.uses_plan(idx) else torch.zeros(3))
            assert torch.equal(got, want) and torch.equal(push.dx(x, got), dmu)
        want, dmu = o_lqr.act(x, None)
        assert torch.allclose(lqr(x), want) and torch.allclose(lqr.dx(x), dmu)

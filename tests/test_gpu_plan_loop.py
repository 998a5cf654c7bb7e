"""klerg_plan_optimize - the optimisation loop of Robot.kldiv_planner decided on the device - against the same
Robot with the loop on the host (two read-backs per iteration): the evals are the same launches on the same
inputs, the decisions the same float32 comparisons, so plans, costs and eval counts must be IDENTICAL, step after
step.  (The host loop itself is checked against the reference's recorded sequences in test_gpu_parity.)"""
import numpy as np
import pytest
import torch

import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import ROBOT_CASES  # noqa: E402
from tests.test_gpu_parity import ct, make_robot  # noqa: E402,F401  (fixture)

pytestmark = pytest.mark.gpu


def _both(host, dev, case):
    """One step of each robot from the same state of torch's global generator (the steps draw from it)."""
    rng = torch.get_rng_state()
    out_h = host.step(case["n"], case["m"], save_update=True)
    after = torch.get_rng_state()
    torch.set_rng_state(rng)
    out_d = dev.step(case["n"], case["m"], save_update=True)
    assert torch.equal(torch.get_rng_state(), after)
    return out_h, out_d


@pytest.mark.parametrize("name", [n for n, c in ROBOT_CASES.items() if not c.get("plot")])
def test_device_loop_equals_host_loop(ct, name):  # noqa: F811
    robots = []
    for device_loop in (False, True):
        r, case = make_robot(ct, name)
        r.device_loop = device_loop
        robots.append(r)
    host, dev = robots
    n_iter = 0
    for k in range(2 * case["steps"]):
        out_h, out_d = _both(host, dev, case)
        for a, b in zip(out_h, out_d):
            np.testing.assert_array_equal(a, b, err_msg=f"{name} step {k}")
        assert torch.equal(host.u, dev.u), f"{name} step {k}: plan"
        assert torch.equal(host.last_plan, dev.last_plan), f"{name} step {k}: planned trajectory"
        assert float(host.last_cost) == float(dev.last_cost), f"{name} step {k}: cost"
        assert host.stats["cost_evals"] == dev.stats["cost_evals"] and host.stats["grad_evals"] == dev.stats["grad_evals"], \
            f"{name} step {k}: eval counts {host.stats} vs {dev.stats}"
        n_iter = dev.stats["grad_evals"]
    assert n_iter >= 2 * case["steps"]  # at least one gradient eval per step was made


def test_device_loop_fixed_lam_and_many_iterations(ct):  # noqa: F811
    """fixed_lam replaces the line search by the window [t_app, t_app + lam); a long inner loop (10 iterations) runs
    into the reference's `break`s - the remaining gated evals must leave the plan alone."""
    for fixed in (True, False):
        robots = []
        for device_loop in (False, True):
            r, case = make_robot(ct, "xyz_small")
            r.device_loop = device_loop
            r.fixed_lam, r.lam = fixed, 3
            r.num_iters_per_step = 10
            robots.append(r)
        host, dev = robots
        for k in range(8):
            _both(host, dev, case)
            assert torch.equal(host.u, dev.u), f"fixed_lam={fixed} step {k}"
            assert float(host.last_cost) == float(dev.last_cost)
            assert host.stats["cost_evals"] == dev.stats["cost_evals"] and host.stats["grad_evals"] == dev.stats["grad_evals"]


def test_plan_optimize_argument_checks(ct):  # noqa: F811
    from control_torch import _cabi as cabi
    lib = cabi.load()
    assert lib.klerg_plan_result_floats(20, 4, 2) == 8 + 40 + 21 * 4
    assert lib.klerg_plan_scratch_bytes(20, 4, 2, 1024) > 7 * 1024 * 4
    with pytest.raises(RuntimeError, match="null"):
        cabi.check(lib.klerg_plan_optimize(None, None, None, None, None, None, None, 20, None, 10, 12, None, None, None, 1e-6,
                                           None, 1.0, None, None, 1, 0, 1, 5, None, None, None, None), "klerg_plan_optimize")

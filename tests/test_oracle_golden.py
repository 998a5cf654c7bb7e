"""Pin the CPU oracle against golden vectors recorded from the live reference.

The golden files were produced by tests/golden/make_golden.py, which imports the
reference's own control_torch package.  Index/sample/selection data must match
bit for bit; fp32 quantities to 2e-6 relative (identical op order, but another
host CPU may vectorise reductions differently).
"""
import os

import numpy as np
import pytest
import torch

from cases import ROBOT_CASES, MixtureTarget, apply_case_flags, robot_kwargs, seed_buffer_states
from oracle import klerg_oracle as ko

RT = 2e-6


def close(a, b, rtol=RT, atol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol + rtol * scale * 1e-1)


@pytest.fixture(scope="module")
def utils(golden_dir):
    return np.load(os.path.join(golden_dir, "utils.npz"))


@pytest.mark.parametrize("D", [2, 3, 4, 6])
def test_pairwise_utils(utils, D):
    t = lambda k: torch.from_numpy(utils[f"D{D}/{k}"])
    traj, samples, std, nu, w = t("traj"), t("samples"), t("std"), t("nu"), t("w")
    explr = torch.arange(D)
    close(ko.footprint_sum(traj, samples, explr, std, nu), utils[f"D{D}/footprint"])
    close(ko.spread_max(traj, samples, explr, std, nu), utils[f"D{D}/spread"])
    g = torch.stack([ko.kl_gradient(x, samples, explr, std, w, nu) for x in traj[:5]])
    close(g, utils[f"D{D}/grad"])
    q = ko.footprint_sum(traj, samples, explr, std, nu)
    close(ko.renormalize(q.clone()), utils[f"D{D}/renorm"])
    close(ko.renormalize_closed_form(q.clone()), utils[f"D{D}/renorm"], rtol=5e-6)
    close(ko.unit_mass(q.clone()), utils[f"D{D}/cost_norm"])


def test_renormalize_floor_and_nan(utils):
    x = torch.from_numpy(utils["floor/x"])
    close(ko.renormalize(x.clone()), utils["floor/renorm"])
    close(ko.renormalize_closed_form(x.clone()), utils["floor/renorm"], rtol=5e-6)
    close(ko.unit_mass(torch.from_numpy(utils["floor/x_nan"]).clone()), utils["floor/cost_norm_nan"])


def test_barrier(utils):
    lim = torch.from_numpy(utils["barrier/lim"])
    xs = torch.from_numpy(utils["barrier/x"])
    bar = ko.OracleBarrier(lim, weight=5.0, power=[4.0] * 6, buff=0.1)
    close(bar.rows(xs), utils["barrier/value"])
    close(torch.stack([bar.grad(x) for x in xs]), utils["barrier/grad"])


@pytest.mark.parametrize("kind,st", [("double", "xyz"), ("speed", "xy"), ("roll", "xyzrpw"), ("single", "xyz")])
def test_dynamics(utils, kind, st):
    x0 = torch.from_numpy(utils[f"dyn_{kind}/x0"])
    u = torch.from_numpy(utils[f"dyn_{kind}/u"])
    env = ko.OracleDynamics(kind, 0.2, x0.clone(), st)
    xs, As, Bs = [env.state.clone()], [], []
    for ut in u:
        a, b = env.get_lin(env.state.clone(), ut)
        As.append(a)
        Bs.append(b)
        xs.append(env.step(ut).clone())
    close(torch.stack(xs), utils[f"dyn_{kind}/traj"])
    close(torch.stack(As), utils[f"dyn_{kind}/A"])
    close(torch.stack(Bs), utils[f"dyn_{kind}/B"])
    if kind == "roll":
        close(env.R, utils["dyn_roll/R_final"])
        env2 = ko.OracleDynamics(kind, 0.2, x0.clone(), st)
        y = torch.stack([env2.step(ut, save=False).clone() for ut in u[:4]])
        close(y, utils["dyn_roll/nosave"])
        close(env2.R, utils["dyn_roll/nosave_R"])


def test_buffer_bit_exact(utils):
    torch.manual_seed(99)
    buf = ko.OracleBuffer(7, 4)
    assert tuple(buf.sample(5).shape) == tuple(utils["buffer/empty_sample_shape"])
    for i, row in enumerate(torch.from_numpy(utils["buffer/seq"])):
        buf.push(row)
        assert np.array_equal(buf.sample(3).numpy(), utils[f"buffer/draw{i}"])
        assert [len(buf), buf.position, int(buf.full)] == utils[f"buffer/len{i}"].tolist()
    assert np.array_equal(buf.get_all().numpy(), utils["buffer/all"])
    assert np.array_equal(buf.get_recent(5).numpy(), utils["buffer/recent5"])
    assert np.array_equal(buf.sample(100).numpy(), utils["buffer/big_draw"])


def build_oracle_robot(name):
    case = ROBOT_CASES[name]
    torch.manual_seed(1234)
    target = MixtureTarget(case["D"], seed=7)
    if case["states"] == "xyzrpw":
        target.mu[:, 3] = target.mu[:, 3] * 0.5 + 3.1
    r = ko.OracleRobot(**robot_kwargs(case, target))
    apply_case_flags(r, case)
    r.test(case["n"])
    for s in seed_buffer_states(r.robot.state, case):
        r.memory_buffer.push(s)
    return r, case


@pytest.mark.parametrize("name", list(ROBOT_CASES))
def test_robot_sequences(golden_dir, name):
    """Whole Robot.step() sequences: same RNG stream, same decisions, same numbers."""
    gold = np.load(os.path.join(golden_dir, f"robot_{name}.npz"))
    r, case = build_oracle_robot(name)
    for k in range(int(gold["n_steps"])):
        pre = f"step{k}/"
        assert np.array_equal(r.u.numpy(), gold[pre + "u_before"]) or np.allclose(r.u.numpy(), gold[pre + "u_before"], rtol=RT, atol=1e-7)
        r.trace = []
        st, vel, ctrl = r.step(case["n"], case["m"], save_update=True)
        # bit-exact: samples (host RNG) and memory-buffer selection
        assert np.array_equal(r._step_inputs["samples"].numpy(), gold[pre + "samples"])
        hist_tol = 0 if k == 0 else 1e-6
        np.testing.assert_allclose(r._step_inputs["hist"].numpy(), gold[pre + "hist"], rtol=hist_tol, atol=hist_tol)
        close(r._step_inputs["p"], gold[pre + "p"])
        close(r._step_inputs["q_base"], gold[pre + "q_base"])
        order = [0 if t["kind"] == "cost" else 1 for t in r.trace]
        assert order == gold[pre + "order"].tolist(), f"control flow diverged at step {k}"
        nc = ng = 0
        for t in r.trace:
            if t["kind"] == "cost":
                close(t["u"], gold[pre + f"cost{nc}/u"], atol=1e-7)
                close(t["cost"], gold[pre + f"cost{nc}/cost"], rtol=1e-5)
                nc += 1
            else:
                close(t["q"], gold[pre + f"grad{ng}/q"])
                close(t["traj"], gold[pre + f"grad{ng}/traj"], atol=1e-7)
                close(t["du"], gold[pre + f"grad{ng}/du"], rtol=1e-5, atol=1e-7)
                close(t["djdlam"], gold[pre + f"grad{ng}/djdlam"], rtol=1e-5, atol=1e-9)
                ng += 1
        close(st, gold[pre + "ret_state"], atol=1e-7)
        close(vel, gold[pre + "ret_vel"], atol=1e-7)
        close(ctrl, gold[pre + "ret_ctrl"], atol=1e-7)
        close(r.u, gold[pre + "u_after"], atol=1e-7)
        close(r.last_plan, gold[pre + "last_plan"], atol=1e-7)
        assert r.memory_buffer.position == int(gold[pre + "buf_pos"])
        if case.get("plot"):
            for i in range(7):
                close(torch.as_tensor(r.plot_data[i]), gold[pre + f"plot{i}"], rtol=1e-5, atol=1e-7)
    close(r.memory_buffer.get_all(), gold["buffer_final"], atol=1e-7)

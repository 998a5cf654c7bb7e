#!/usr/bin/env python
"""Generate golden vectors by running the LIVE reference (read-only /root/reference).

The reference ships no tests or fixtures, so parity is
pinned against outputs of the reference itself, produced in the authoring
container (torch CPU, fp32) by this script:

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Nothing here is imported by the product.  The reference is imported in place
(never copied): ``/root/reference/franka_test/scripts`` is put on ``sys.path``
together with a 2-line ``termcolor`` stand-in (klerg.py:7 imports it; it is not
installed in this image).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SCRIPTS = "/root/reference/franka_test/scripts"


def import_reference():
    shim = types.ModuleType("termcolor")
    shim.cprint = lambda *a, **k: None
    shim.colored = lambda s, *a, **k: s
    sys.modules.setdefault("termcolor", shim)
    if REF_SCRIPTS not in sys.path:
        sys.path.insert(0, REF_SCRIPTS)
    import control_torch.klerg as rk  # noqa
    import control_torch.klerg_utils as ru  # noqa
    import control_torch.barrier as rb  # noqa
    import control_torch.dynamics as rd  # noqa
    import control_torch.memory_buffer as rm  # noqa
    return rk, ru, rb, rd, rm


sys.path.insert(0, HERE)
from cases import MixtureTarget, ROBOT_CASES, apply_case_flags, robot_kwargs, seed_buffer_states  # noqa: E402


def record_robot_case(rk, name, case):
    torch.manual_seed(1234)
    target = MixtureTarget(case["D"], seed=7)
    if case["states"] == "xyzrpw":
        target.mu[:, 3] = target.mu[:, 3] * 0.5 + 3.1
    r = rk.Robot(**robot_kwargs(case, target))
    apply_case_flags(r, case)
    out = {}
    log = []

    orig_cost, orig_back = r.get_cost, r.backward

    def cost_hook(samples, p, q_base, hist, u_test, u_def, receding_barrier=False):
        rec = dict(samples=samples.clone(), p=p.clone(), q_base=q_base.clone(), hist=hist.clone(), u=u_test.clone(), x0=r.robot.state.clone())
        c = orig_cost(samples, p, q_base, hist, u_test, u_def, receding_barrier)
        rec["cost"] = c.clone()
        log.append(("cost", rec))
        return c

    def back_hook(samples, p, q, nu, grad_list, traj_list):
        du, dj = orig_back(samples, p, q, nu, grad_list, traj_list)
        A = torch.stack([g[0] for g in grad_list])
        dbarr = torch.stack([g[2] for g in grad_list])
        log.append(("grad", dict(u=r.u.clone(), q=q.clone(), traj=traj_list.clone(), du=du.clone(), djdlam=dj.clone(), A=A, dbarr=dbarr)))
        return du, dj

    r.get_cost, r.backward = cost_hook, back_hook
    r.test(case["n"])
    # seed the buffer with a short random walk so the first steps have history
    for s in seed_buffer_states(r.robot.state, case):
        r.memory_buffer.push(s)
    for k in range(case["steps"]):
        log.clear()
        u_before = r.u.clone()
        state_before = r.robot.state.clone()
        buf_len = len(r.memory_buffer)
        st, vel, ctrl = r.step(case["n"], case["m"], save_update=True)
        pre = f"step{k}/"
        out[pre + "u_before"] = u_before.numpy()
        out[pre + "state_before"] = state_before.numpy()
        out[pre + "buf_len_before"] = np.array(buf_len)
        out[pre + "ret_state"], out[pre + "ret_vel"], out[pre + "ret_ctrl"] = st, vel, ctrl
        out[pre + "u_after"] = r.u.numpy().copy()
        out[pre + "last_plan"] = r.last_plan.numpy().copy()
        out[pre + "state_after"] = r.robot.state.numpy().copy()
        out[pre + "buf_pos"] = np.array(r.memory_buffer.position)
        first = log[0][1]
        for key in ("samples", "p", "q_base", "hist"):
            out[pre + key] = first[key].numpy()
        nc = ng = 0
        order = []
        for kind, rec in log:
            if kind == "cost":
                out[pre + f"cost{nc}/u"] = rec["u"].numpy()
                out[pre + f"cost{nc}/cost"] = rec["cost"].numpy()
                order.append(0)
                nc += 1
            else:
                for key, v in rec.items():
                    out[pre + f"grad{ng}/{key}"] = v.numpy()
                order.append(1)
                ng += 1
        out[pre + "order"] = np.array(order)
        if case.get("plot"):
            for i, pd in enumerate(r.plot_data):
                out[pre + f"plot{i}"] = torch.as_tensor(pd).numpy().copy()
    out["buffer_final"] = r.memory_buffer.get_all().numpy()
    out["n_steps"] = np.array(case["steps"])
    np.savez_compressed(os.path.join(HERE, f"robot_{name}.npz"), **out)
    print(name, "->", len(out), "arrays")


def record_utils(ru, rb, rd, rm):
    out = {}
    g = torch.Generator().manual_seed(11)
    for D, S, T, N in [(2, 4, 37, 301), (3, 6, 64, 257), (6, 12, 23, 130), (4, 4, 19, 200)]:
        tag = f"D{D}/"
        traj = torch.rand(T, S, generator=g) * 2 - 1
        samples = torch.rand(N, D, generator=g) * 2.3 - 1.15
        explr = torch.arange(D)
        std = (torch.rand(D, generator=g) * 0.1 + 0.03) * torch.tensor([1.0, -1.0] * (D // 2) + [1.0] * (D % 2))
        nu = torch.tensor([1.0]) if D != 3 else torch.tensor(7.0)
        w = torch.rand(N, generator=g) + 0.1
        out[tag + "traj"], out[tag + "samples"], out[tag + "std"], out[tag + "nu"], out[tag + "w"] = (
            traj.numpy(), samples.numpy(), std.numpy(), nu.numpy(), w.numpy())
        out[tag + "footprint"] = ru.traj_footprint_vec(traj, samples, explr, std, nu).numpy()
        out[tag + "spread"] = ru.traj_spread_vec(traj, samples, explr, std, nu).numpy()
        out[tag + "grad"] = torch.stack([ru.kldiv_grad_vec(x, samples, explr, std, w, nu) for x in traj[:5]]).numpy()
        q = ru.traj_footprint_vec(traj, samples, explr, std, nu)
        out[tag + "renorm"] = ru.renormalize(q.clone()).numpy()
        out[tag + "cost_norm"] = ru.cost_norm(q.clone()).numpy()
    # renormalize floor behaviour: wide dynamic range + NaN handling in cost_norm
    x = torch.exp(torch.linspace(-30, 0, 500))
    out["floor/x"] = x.numpy()
    out["floor/renorm"] = ru.renormalize(x.clone()).numpy()
    xn = x.clone()
    xn[::50] = float("nan")
    out["floor/x_nan"] = xn.numpy()
    out["floor/cost_norm_nan"] = ru.cost_norm(xn.clone()).numpy()

    # barrier
    lim = torch.tensor([[-1.0, 1.0], [-1.0, 1.0], [-0.75, 0.75], [-1.25, 1.25], [-1.25, 1.25], [-0.5, 0.5]])
    bar = rb.BarrierFunction(b_lim=lim, barr_weight=5.0, b_buff=0.1, power=[4.0] * 6)
    xs = torch.rand(40, 6, generator=g) * 3.2 - 1.6
    xs[0] = torch.tensor([0.9, -0.9, 0.65, 1.15, -1.15, 0.4])  # exactly on the shrunk walls
    out["barrier/lim"], out["barrier/x"] = lim.numpy(), xs.numpy()
    out["barrier/value"] = bar(xs).numpy()
    out["barrier/grad"] = torch.stack([bar.dbarr(x) for x in xs]).numpy()

    # dynamics rollouts
    for kind, cls, S, A, st in [("double", rd.DoubleIntegratorEnv, 6, 3, "xyz"), ("speed", rd.DoubleIntegratorSpeedEnv, 6, 2, "xy"),
                                ("roll", rd.DoubleIntegratorRollEnv, 12, 6, "xyzrpw"), ("single", rd.SingleIntegratorEnv, 3, 3, "xyz")]:
        x0 = torch.rand(S, generator=g) * 0.6 - 0.3
        if kind == "roll":
            x0[3] += 3.0
        if kind == "speed":
            x0[4:] = x0[2:4].abs()
        env = cls(dt=0.2, x0=x0.clone(), states=st)
        u = torch.rand(25, A, generator=g) * 2 - 1
        xs_, As, Bs = [env.state.clone()], [], []
        for ut in u:
            a_, b_ = env.get_lin(env.state.clone(), ut)
            As.append(a_.clone())
            Bs.append(b_.clone())
            xs_.append(env.step(ut).clone())
        out[f"dyn_{kind}/x0"], out[f"dyn_{kind}/u"] = x0.numpy(), u.numpy()
        out[f"dyn_{kind}/traj"] = torch.stack(xs_).numpy()
        out[f"dyn_{kind}/A"], out[f"dyn_{kind}/B"] = torch.stack(As).numpy(), torch.stack(Bs).numpy()
        if kind == "roll":
            out["dyn_roll/R_final"] = env.R.numpy()
            # save=False still advances R
            env2 = cls(dt=0.2, x0=x0.clone(), states=st)
            y = [env2.step(ut, save=False).clone() for ut in u[:4]]
            out["dyn_roll/nosave"] = torch.stack(y).numpy()
            out["dyn_roll/nosave_R"] = env2.R.numpy()

    # memory buffer: indices must be bit exact
    torch.manual_seed(99)
    buf = rm.MemoryBuffer_torch(7, 4, dtype=torch.float32)
    seq = torch.arange(44, dtype=torch.float32).reshape(11, 4)
    out["buffer/empty_sample_shape"] = np.array(buf.sample(5).shape)
    for i, row in enumerate(seq):
        buf.push(row)
        out[f"buffer/draw{i}"] = buf.sample(3).numpy()
        out[f"buffer/len{i}"] = np.array([len(buf), buf.position, int(buf.full_buffer)])
    out["buffer/all"] = buf.get_all().numpy()
    out["buffer/recent5"] = buf.get_recent(5).numpy()
    out["buffer/big_draw"] = buf.sample(100).numpy()
    out["buffer/seq"] = seq.numpy()
    np.savez_compressed(os.path.join(HERE, "utils.npz"), **out)
    print("utils ->", len(out), "arrays")


def record_barrier_variants(rb):
    """TiltBarrierFunction and VelocityBarrier (barrier.py:95-144, 162-205): the two variants the reference keeps
    beside BarrierFunction -> barrier_variants.npz."""
    from control_torch.klerg_utils import Lambda
    from franka.franka_utils import ws_conversion
    out = {}
    g = torch.Generator().manual_seed(23)
    lim = torch.tensor([[-1.0, 1.0], [-1.0, 1.0], [-0.75, 0.75], [2.39, 3.89], [-0.75, 0.75], [-2.0, 2.0]] + [[-1.25, 1.25]] * 6)
    xs = torch.rand(48, 12, generator=g) * 2.4 - 1.2
    xs[:, 3] = 2.0 + 2.2 * torch.rand(48, generator=g)
    xs[:, 5] = 5.0 * torch.rand(48, generator=g) - 2.5
    out["tilt/lim"], out["tilt/x"] = lim.numpy(), xs.numpy()
    robot_rpw = torch.tensor([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]])
    tray_rpw = torch.tensor([[2.39, 3.89], [-0.75, 0.75], [-2.0, 2.0]])
    for tag, fn, x_in in (("plain", None, xs), ("mapped", Lambda(ws_conversion, (robot_rpw, tray_rpw)), None)):
        if x_in is None:  # roll / pitch / yaw columns in the unit "robot" range the map expects
            x_in = xs.clone()
            x_in[:, 3:6] = torch.rand(48, 3, generator=g) * 1.3 - 0.15
            out["tilt/x_mapped"] = x_in.numpy()
        other = rb.BarrierFunction(b_lim=lim, barr_weight=5.0, b_buff=0.1, power=[4.0] * 12)
        bar = rb.TiltBarrierFunction(other, "xyzrpw", tilt_lim=2.45, rot_to_angles_fn=fn)
        out[f"tilt/{tag}/w_b_lim"] = bar.w_b_lim.numpy().copy()
        out[f"tilt/{tag}/value"] = bar(x_in).numpy()
        out[f"tilt/{tag}/grad"] = torch.stack([bar.dbarr(x) for x in x_in]).numpy()
        out[f"tilt/{tag}/w_lim_after"] = other.b_lim[5].numpy().copy()
    vb = rb.VelocityBarrier("xyzXYZ", b_lim=0.1, power=4, barr_weight=100.0)
    x_old = torch.rand(40, 6, generator=g) * 2 - 1
    x_new = x_old + 0.5 * (torch.rand(40, 6, generator=g) - 0.5)
    x_new[3] = x_old[3]
    x_new[4, 3:] = x_old[4, 3:] + torch.tensor([0.1, -0.1, 0.05])  # exactly on / inside the band
    out["vel/x_old"], out["vel/x_new"] = x_old.numpy(), x_new.numpy()
    out["vel/value"] = torch.stack([vb.barr(a, b) for a, b in zip(x_new, x_old)]).numpy()
    out["vel/grad"] = torch.stack([vb.dbarr(a, b) for a, b in zip(x_new, x_old)]).numpy()
    out["vel/call"] = vb(x_old, x_new).numpy()
    np.savez_compressed(os.path.join(HERE, "barrier_variants.npz"), **out)
    print("barrier variants ->", len(out), "arrays")


def main():
    rk, ru, rb, rd, rm = import_reference()
    torch.set_num_threads(1)  # fixed reduction order for reproducible fixtures
    only = sys.argv[1:]  # optional: names of the robot cases to (re)record; default = everything
    if not only:
        record_utils(ru, rb, rd, rm)
    if not only or "barrier_variants" in only:
        record_barrier_variants(rb)
    for name, case in ROBOT_CASES.items():
        if not only or name in only:
            record_robot_case(rk, name, case)


if __name__ == "__main__":
    main()

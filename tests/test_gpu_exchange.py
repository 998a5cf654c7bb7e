"""The cross-rank exchange protocol of the fused evals on a ONE-GPU box: two emulated ranks, each owning half of the
workspace samples and its own mailbox, run as ONE cooperative launch whose halves act as the two ranks
(``engine.emulated_pair``).  The device code is the code the sharded launches run; only the transport differs (the
"peer" mailbox is local memory instead of an NVLink mapping).  Checked: both ranks hold bit-identical results (they
drive identical host control flow), the results match the single-rank eval of the whole workspace to 1e-5 of the
largest entry (the sums are taken in a different order), the tag / parity / counter bookkeeping survives many
consecutive evals of alternating kinds, and no wait timed out.  The real 2/4/8-GPU runs: tests/mp/, profiles/."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import workloads as wl  # noqa: E402


def _setup(name, n_total, m, H=None, seed=0):
    from control_torch import engine
    from control_torch.klerg import Robot
    from control_torch.planner import PlannerContext
    w = wl.WORKLOADS[name]
    H = H or w["H"]
    lims = [wl.LIMS[s] for s in w["states"]]
    D = len(lims)
    dev = torch.device("cuda")
    target = wl.make_target("gmm", lims, seed=1, device=dev)
    kw = wl.robot_kwargs(name, target, n_samples=n_total, horizon=H, cap=max(m, 8))
    torch.manual_seed(seed)
    probe = Robot(**kw)
    g = torch.Generator().manual_seed(seed)
    lo = torch.tensor([a for a, _ in lims]) * 1.15
    hi = torch.tensor([b for _, b in lims]) * 1.15
    smp = (lo + torch.rand(n_total, D, generator=g) * (hi - lo)).to(dev)
    p_raw = target.pdf_torch(smp).contiguous()
    p_all, p_stats, _ = engine.target_weight(2, smp, lo.tolist(), hi.tolist(), None, p_raw, n_total, 1.0, True, engine.SINGLE)
    hist = wl.random_walk_history(name, m, seed=seed).to(dev)
    x0 = torch.tensor(kw["x0"], dtype=torch.float32, device=dev)

    def make(group, a, b):
        ctx = PlannerContext(probe.planner.spec, probe.barrier.spec(), probe.explr_locs.tolist(), H,
                             torch.diagonal(probe.R_inv).tolist(), probe.control_lim[:, 0].tolist(),
                             probe.control_lim[:, 1].tolist(), alpha=1.0, group=group)
        ctx.set_samples(smp[a:b].contiguous(), probe.std.tolist(), 1.0, n_total=n_total)
        ctx.set_state(x0)
        ctx.set_target(p_all[a:b].contiguous(), p_stats)
        ctx.set_history(hist)
        return ctx

    g0, g1 = engine.EmulatedShardGroup.make_pair(dev)
    r0, r1 = make(g0, *g0.shard_bounds(n_total)), make(g1, *g1.shard_bounds(n_total))
    single = make(engine.SINGLE, 0, n_total)
    return dict(engine=engine, ranks=(r0, r1), single=single, D=D, H=H, dev=dev, lims=lims, smp=smp, n=n_total)


def _rel(x, y):
    x, y = x.double().cpu().numpy(), y.double().cpu().numpy()
    return float(np.abs(x - y).max() / (np.abs(y).max() + 1e-30))


@pytest.mark.parametrize("name,n_total,H", [("c2", 40_003, 50), ("c4", 30_001, 50), ("c4", 9_001, 23), ("c1", 2_000, 20)])
def test_two_emulated_ranks_match_single_rank(name, n_total, H):
    s = _setup(name, n_total, 400, H=H)
    eng, (r0, r1), single = s["engine"], s["ranks"], s["single"]
    U = wl.random_controls((6, H, s["D"]), seed=3).to(s["dev"])
    for it in range(5):  # consecutive evals of both kinds: slot parities, tags and counters must keep working
        nb = 1 + it  # 1..5 candidates: one-hop and two-stage all-reduce
        c0, c1 = eng.emulated_pair(lambda: r0.costs(U[:nb], view=True), lambda: r1.costs(U[:nb], view=True))
        want_c = single.costs(U[:nb])
        assert torch.equal(c0, c1), f"iter {it}: the two ranks hold different costs"
        assert _rel(c0, want_c) < 1e-5, (it, c0.tolist(), want_c.tolist())
        ga, gb = eng.emulated_pair(lambda: r0.gradient(U[it], keep=True, want_cost=(it % 2 == 1)),
                                   lambda: r1.gradient(U[it], keep=True, want_cost=(it % 2 == 1)))
        gs = single.gradient(U[it], keep=True, want_cost=(it % 2 == 1))
        keys = ["du", "djdlam", "u_star", "dgdx", "totals"] + (["cost"] if it % 2 == 1 else [])
        for k in keys:
            assert torch.equal(ga[k], gb[k]), f"iter {it}: {k} differs between the ranks"
            assert _rel(ga[k], gs[k]) < 1e-5, (it, k, _rel(ga[k], gs[k]))
        assert float(ga["host_pack"][-1]) == 0.0
    torch.cuda.synchronize()
    assert not eng.fused_fault()


def test_emulated_ranks_with_belief_targets():
    """K targets in one launch: one gather exchange per target, alternating slot parity inside the kernel."""
    s = _setup("c2", 30_000, 300, H=24)
    eng, (r0, r1), single = s["engine"], s["ranks"], s["single"]
    K, n = 3, s["n"]
    P_all = torch.stack([wl.make_target("gmm", s["lims"], seed=30 + k, device=s["dev"]).pdf_torch(s["smp"]) for k in range(K)])
    stats = torch.stack([P_all[k].double().sum().reshape(1) for k in range(K)])
    for ctx, grp in ((r0, r0.group), (r1, r1.group)):
        a, b = grp.shard_bounds(n)
        ctx.set_targets(P_all[:, a:b].contiguous(), stats)
        ctx.targets_path = "fused"
    single.set_targets(P_all.contiguous(), stats)
    single.targets_path = "fused"
    u = wl.random_controls((24, s["D"]), seed=5).to(s["dev"])
    for _ in range(2):
        ga, gb = eng.emulated_pair(lambda: r0.gradient_targets(u), lambda: r1.gradient_targets(u))
        gs = single.gradient_targets(u)
        for k in ("du", "djdlam", "u_star", "dgdx"):
            assert torch.equal(ga[k], gb[k]), k
            assert _rel(ga[k], gs[k]) < 1e-5, (k, _rel(ga[k], gs[k]))
    torch.cuda.synchronize()
    assert not eng.fused_fault()


def test_back_to_back_evals_with_overlap_enabled():
    """Independent evals launched back to back with KLERG_OPT_EVAL_OVERLAP (the next eval's CTAs start while the
    previous eval's finisher CTA is still in its tail) give the same bits as the serialised launches."""
    s = _setup("c2", 100_000, 500, H=50)
    eng, single = s["engine"], s["single"]
    U = wl.random_controls((8, 50, s["D"]), seed=9).to(s["dev"])
    want = []
    for i in range(8):
        g = single.gradient(U[i])
        want.append({k: g[k].clone() for k in ("du", "djdlam", "u_star", "dgdx")})
    want_c = [single.costs(U[i:i + 3]).clone() for i in range(4)]
    torch.cuda.synchronize()
    eng.set_eval_overlap(True)
    try:
        outs = []
        bufs = [eng.EvalBuffers(50, single.dyn.S, single.dyn.A, single.packed.shape[1], s["dev"]) for _ in range(8)]
        for rep in range(3):
            outs.clear()
            for i in range(8):
                single.buf = bufs[i]  # every eval of the batch writes its own outputs
                g = single.gradient(U[i])
                outs.append({k: g[k] for k in ("du", "djdlam", "u_star", "dgdx")})
            torch.cuda.synchronize()
            for i in range(8):
                for k in want[i]:
                    assert torch.equal(outs[i][k], want[i][k]), (rep, i, k)
        got_c = []
        for i in range(4):
            single.buf = bufs[i]
            got_c.append(single.costs(U[i:i + 3]))
        torch.cuda.synchronize()
        for i in range(4):
            assert torch.equal(got_c[i], want_c[i]), i
    finally:
        eng.set_eval_overlap(False)
    assert not eng.fused_fault()

"""Device-resident evaluation context of the KL-ergodic planner.

One ``PlannerContext`` holds everything an *eval* needs on the GPU - packed
workspace samples, target density p, history footprint q_base, the planner
model and barrier - and exposes the two batched operations the reference
performs through Python loops:

* ``costs(U)``      = ``Robot.get_cost`` for B candidate control sequences
                      (klerg.py:686-710): rollout, barrier, footprint over the
                      post-step states, renormalise, KL(p||q);
* ``gradient(u)``   = ``Robot.forward`` + footprint + ``Robot.backward``
                      (klerg.py:409-450): rollout with linearisation, footprint
                      over the pre-step states, importance ratio, dgdx for all
                      H steps, adjoint sweep -> du, djdlam, u*.

With a ``ShardGroup`` of world > 1 each rank holds a contiguous slice of the
samples; the only traffic is one all-gather of [G,2] totals after the forward
pass and one of [H*D+2] partials after the gradient pass.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi as cabi
from . import engine


class PlannerContext:
    def __init__(self, dyn, barrier_spec, explr_locs, horizon, rinv_diag, ctrl_lo, ctrl_hi, alpha=1.0,
                 group=engine.SINGLE, floor=engine.FLOOR, fused=True):
        cabi.require_cuda()
        self.dyn = dyn
        self.bar = barrier_spec
        self.explr_locs = [int(i) for i in explr_locs]
        self.H = int(horizon)
        self.rinv = [float(v) for v in rinv_diag]
        self.ctrl_lo = [float(v) for v in ctrl_lo]
        self.ctrl_hi = [float(v) for v in ctrl_hi]
        self.alpha = float(alpha)
        self.group = group
        self.floor = floor
        self.spec = None
        self.n = 0
        self.evals = dict(cost=0, grad=0, fwd_pairs=0, grad_pairs=0)
        self.fused = fused
        self.peers = group.peers() if fused else None
        self._peers_ref = C.byref(self.peers) if self.peers is not None else None
        self.R0 = None
        self._rinv_c, self._lo_c, self._hi_c = cabi.farr(self.rinv), cabi.farr(self.ctrl_lo), cabi.farr(self.ctrl_hi)
        self.buf = None

    # -- per-step inputs ---------------------------------------------------------
    def set_samples(self, samples_dev, scale, nu=1.0, n_total=None):
        """samples_dev [n_local, D] raw (this rank's slice); scale = std as the reference uses it.  ``n_total`` = samples
        over all ranks (a sharded context sizes its fused grids from the largest shard, identically on every rank)."""
        D = samples_dev.shape[1]
        if self.peers is not None:
            if n_total is None:
                n_local = torch.tensor([samples_dev.shape[0]], device=samples_dev.device)
                n_total = int(self.group.gather_blocks(n_local).sum().item())
            self.peers.n_max = self.group.max_shard(n_total)
        self.spec = cabi.kernel_spec(D, self.dyn.S, self.explr_locs, [float(s) for s in scale], nu)
        self.samples = samples_dev
        self.n = samples_dev.shape[0]
        self.packed = engine.pack_samples(self.spec, samples_dev)
        ld = self.packed.shape[1]
        if self.buf is None or self.buf.ld != ld or self.buf.H != self.H or self.buf.A != self.dyn.A or self.buf.S != self.dyn.S:
            self.buf = engine.EvalBuffers(self.H, self.dyn.S, self.dyn.A, ld, samples_dev.device)

    def set_target(self, p, p_stats):
        """p [n_local] (or already padded to the sample stride ld), p_stats [1] = global sum of p."""
        ld = self.packed.shape[1]
        if p.numel() < ld or p.data_ptr() % 16:  # the fused evals copy whole tile rows with TMA: rows are padded to the
            pad = torch.zeros(ld, dtype=torch.float32, device=p.device)  # sample stride and 16-byte aligned
            pad[: p.numel()] = p
            p = pad
        self.p, self.p_stats = p, p_stats

    def set_history(self, hist_dev):
        """q_base = footprint of the drawn history rows (zeros when empty, klerg.py:496-498)."""
        out, _ = engine.footprint(self.spec, 0, hist_dev, self.packed, self.n)
        self.q_base = out[0]
        return self.q_base

    def set_state(self, x0_dev):
        self.x0 = x0_dev.contiguous()

    # -- evals -------------------------------------------------------------------
    def costs(self, U, view=False):
        """U [B,H,A] on the device -> cost [B] (KL + barrier), all on the device.  ``view=True`` (planner): for
        B <= 8 the result is a view of ``buf.cost_pack``, valid until the next call."""
        if U.dim() == 2:
            U = U.unsqueeze(0)
        B = U.shape[0]
        if self.fused:
            U = U.contiguous()
            mg = self.buf.max_g
            if B <= mg:  # one launch (a line search): costs and the fault word share one buffer -> a single D2H
                pack = self.buf.cost_pack
                engine.eval_costs(self.spec, self.dyn, self.bar, self._peers_ref, self.x0, self.R0, U, self.packed, self.n,
                                  self.q_base, self.p, self.p_stats, self.buf.v_costs, pack[:B], self.floor, fault=pack[mg:])
                cost = pack[:B] if view else pack[:B].clone()
            elif self.peers is None and self.batch_costs:
                # any number of candidates in ONE launch (BASELINE config 3)
                cost, _ = engine.eval_costs_batch(self.spec, self.dyn, self.bar, self.x0, self.R0, U, self.packed, self.n,
                                                  self.q_base, self.p, self.p_stats, self.floor)
            else:
                cost = torch.empty(B, dtype=torch.float32, device=U.device)
                for b0 in range(0, B, mg):
                    b1 = min(B, b0 + mg)
                    engine.eval_costs(self.spec, self.dyn, self.bar, self._peers_ref, self.x0, self.R0, U[b0:b1], self.packed,
                                      self.n, self.q_base, self.p, self.p_stats, self.buf.v_costs, cost[b0:b1], self.floor)
            self.evals["cost"] += B
            self.evals["fwd_pairs"] += B * self.H * self.n
            return cost
        ro = engine.rollout(self.dyn, self.bar, self.x0, U)
        v, totals = engine.footprint(self.spec, 0, ro["traj"][:, 1:, :], self.packed, self.n, add_in=self.q_base)
        totals_w = self.group.gather_blocks(totals)
        self.evals["cost"] += B
        self.evals["fwd_pairs"] += B * self.H * self.n
        return engine.kl_cost(v, self.n, totals_w, self.p, self.p_stats, ro["barrier"], self.group, self.floor)

    def gradient(self, u, keep=False, want_cost=False, policy=None):
        """u [H,A] on the device -> dict(du, djdlam, u_star, dgdx, ...) on the device.  ``policy``: a
        cabi.PolicySpec of a state-feedback default policy (BarrierPush / LQR): the controls are then what the policy
        applies along the closed loop (returned as ``u_eff``) and the adjoint carries dmu/dx."""
        if policy is not None:
            return self._gradient_feedback(u, policy, keep)
        if self.fused:
            o = engine.eval_gradient(self.spec, self.dyn, self.bar, self._peers_ref, self.x0, self.R0, u.reshape(self.H, -1).contiguous(),
                                     self.packed, self.n, self.q_base, self.p, self.p_stats, self._rinv_c, self.alpha,
                                     self._lo_c, self._hi_c, self.buf.next_set(), self.floor, want_cost=want_cost)
            self.evals["grad"] += 1
            self.evals["fwd_pairs"] += self.H * self.n
            self.evals["grad_pairs"] += self.H * self.n
            out = dict(du=o["du"], djdlam=o["djdlam"], u_star=o["u_star"], dgdx=o["dgdx"], traj=o["traj"][: self.H],
                       host_pack=o["host_pack"])
            if want_cost:
                out.update(kl_parts=o["kl"].unsqueeze(0), cost=o["cost"])
            if keep:
                out.update(v=o["v"], totals=o["totals"].unsqueeze(0))
            return out
        ro = engine.rollout(self.dyn, self.bar, self.x0, u, want_lin=True)
        traj = ro["traj"][0]
        pre = traj[: self.H]
        v, totals = engine.footprint(self.spec, 0, pre, self.packed, self.n, add_in=self.q_base)
        totals_w = self.group.gather_blocks(totals)
        gpart, klpart = engine.kl_gradient_fused(self.spec, pre, self.packed, self.n, v[0], totals_w, self.p, self.floor)
        if self.group.world > 1:
            packed = self.group.gather_blocks(torch.cat([gpart.reshape(-1), klpart]))
            gparts = packed[:, :-2].reshape(self.group.world, self.H, self.spec.D).contiguous()
            klparts = packed[:, -2:]
        else:
            gparts, klparts = gpart.unsqueeze(0), klpart.unsqueeze(0)
        P = ro["P"][0] if ro["P"] is not None else None
        dgdx, du, dj, ustar = engine.adjoint(self.dyn, self.spec, gparts, ro["dbarr"][0], P, traj,
                                             u.reshape(self.H, -1), self.rinv, self.alpha, self.ctrl_lo, self.ctrl_hi)
        self.evals["grad"] += 1
        self.evals["fwd_pairs"] += self.H * self.n
        self.evals["grad_pairs"] += self.H * self.n
        out = dict(du=du, djdlam=dj, u_star=ustar, dgdx=dgdx, traj=pre, kl_parts=klparts)
        if keep:
            out.update(v=v[0], totals=totals_w)
        return out

    def _gradient_feedback(self, u, policy, keep):
        """Robot.forward + backward (klerg.py:409-450) for a policy with dmu/dx != 0: the closed loop fixes the controls
        (one serial warp), the pair passes run open-loop on them, the adjoint uses A_t + B dmudx_t."""
        H = self.H
        u_eff, dmudx = engine.policy_rollout(self.dyn, policy, self.x0, self.R0, u.reshape(H, -1) if u is not None else None, H)
        ro = engine.rollout(self.dyn, self.bar, self.x0, u_eff, R0=self.R0, want_lin=True)
        traj = ro["traj"][0]
        pre = traj[:H]
        v, totals = engine.footprint(self.spec, 0, pre, self.packed, self.n, add_in=self.q_base)
        totals_w = self.group.gather_blocks(totals)
        gpart, klpart = engine.kl_gradient_fused(self.spec, pre, self.packed, self.n, v[0], totals_w, self.p, self.floor)
        if self.group.world > 1:
            packed = self.group.gather_blocks(torch.cat([gpart.reshape(-1), klpart]))
            gparts = packed[:, :-2].reshape(self.group.world, H, self.spec.D).contiguous()
        else:
            gparts = gpart.unsqueeze(0)
        P = ro["P"][0] if ro["P"] is not None else None
        dgdx, _, _, _ = engine.adjoint(self.dyn, self.spec, gparts, ro["dbarr"][0], P, traj, u_eff, self.rinv, self.alpha,
                                       self.ctrl_lo, self.ctrl_hi)  # assembles dgdx [H,S] from the partials
        du, dj, ustar = engine.adjoint_policy(self.dyn, dgdx, ro["dbarr"][0], P, dmudx, u_eff, self.rinv, self.alpha,
                                              self.ctrl_lo, self.ctrl_hi)
        self.evals["grad"] += 1
        self.evals["fwd_pairs"] += H * self.n
        self.evals["grad_pairs"] += H * self.n
        out = dict(du=du, djdlam=dj, u_star=ustar, dgdx=dgdx, traj=pre, u_eff=u_eff, dmudx=dmudx)
        if keep:
            out.update(v=v[0], totals=totals_w)
        return out

    def optimize(self, u, num_iters, fixed_lam=False, lam=1):
        """The whole optimisation loop of kldiv_planner (klerg.py:505-576) without host round trips: ``u`` [H,A] on the
        device is replaced by the optimised plan; returns the packed result (device) - see engine.plan_optimize."""
        if not self.fused:
            raise RuntimeError("the device-resident planner loop needs the fused evals")
        return engine.plan_optimize(self.spec, self.dyn, self.bar, self._peers_ref, self.x0, self.R0, u, self.packed, self.n,
                                    self.q_base, self.p, self.p_stats, self._rinv_c, self.alpha, self._lo_c, self._hi_c,
                                    num_iters, fixed_lam, lam, self.buf, self.floor)

    # -- K belief targets over one workspace (fingerprint test mode, BASELINE config 5) ------------------
    def set_targets(self, P, P_stats):
        """P [K, n_local] target densities p_k on the device, P_stats [K, 1] = their global sums."""
        K, n = P.shape
        ld = self.packed.shape[1]
        self.P = torch.zeros((K, ld), dtype=torch.float32, device=P.device)  # rows padded to the sample stride
        self.P[:, :n] = P
        self.P_stats = P_stats.reshape(K).contiguous()

    def costs_targets(self, U):
        """cost[k, b] = KL(p_k || q_b) + barrier_b for every target k and candidate b."""
        saved = (self.p, self.p_stats)
        out = []
        try:
            for k in range(self.P.shape[0]):
                self.set_target(self.P[k, : self.n].contiguous(), self.P_stats[k: k + 1])
                out.append(self.costs(U).clone())
        finally:
            self.p, self.p_stats = saved
        return torch.stack(out)

    def check_fault(self):
        """Raise if an in-kernel wait of any earlier eval of this context's stream timed out (small D2H reads)."""
        if engine.fused_fault() or engine.targets_gradient_fault():
            raise RuntimeError("a fused eval or the targets contraction gave up an in-kernel wait (lost peer or a launch "
                               "that was not co-resident); the results since the last check are void")

    def gradient_targets(self, u, check=True):
        out = self._gradient_targets(u)
        if check:  # the caller reads the results next anyway; ``check=False`` keeps the call asynchronous
            self.check_fault()
        return out

    def _gradient_targets(self, u):
        """Per-target gradient eval: dict of [K, ...] stacked du, djdlam, u_star, dgdx.  Rollout, forward pair pass and
        q are shared by the targets.  Two implementations (``targets_path``): "fused" - one launch, the gradient pair
        pass and the adjoint run per target (psi recomputed per target); "tensor" - psi once per state-sample pair and
        the sum over the samples as a tensor-core contraction for all targets (klerg_kl_gradient_targets), the default
        ("auto") for >= 4 targets with H <= 64 (2.5x faster at BASELINE config 5); sharded contexts all-gather the
        totals and the per-target partials in rank order."""
        K = self.P.shape[0]
        u = u.reshape(self.H, -1).contiguous()
        if self.targets_path == "tensor" or (self.targets_path == "auto" and self._tensor_targets_ok(K)):
            return self._gradient_targets_tensor(u)
        if not self.fused:
            saved, acc = (self.p, self.p_stats), {k: [] for k in ("du", "djdlam", "u_star", "dgdx")}
            try:
                for k in range(K):
                    self.set_target(self.P[k, : self.n].contiguous(), self.P_stats[k: k + 1])
                    g = self.gradient(u)
                    for key in acc:
                        acc[key].append(g[key].clone())
            finally:
                self.p, self.p_stats = saved
            return {k: torch.stack(v) for k, v in acc.items()}
        outs = []
        for k0 in range(0, K, 32):
            k1 = min(K, k0 + 32)
            outs.append(engine.eval_gradient_targets(self.spec, self.dyn, self.bar, self._peers_ref, self.x0, self.R0, u,
                                                     self.packed, self.n, self.q_base, self.P[k0:k1], self.P_stats[k0:k1],
                                                     self._rinv_c, self.alpha, self._lo_c, self._hi_c, self.buf.sets[0]["v"],
                                                     self.floor))
        self.evals["grad"] += K
        self.evals["fwd_pairs"] += self.H * self.n * len(outs)
        self.evals["grad_pairs"] += self.H * self.n * K
        if len(outs) == 1:
            return {k: outs[0][k] for k in ("du", "djdlam", "u_star", "dgdx")}
        return {k: torch.cat([o[k] for o in outs]) for k in ("du", "djdlam", "u_star", "dgdx")}

    batch_costs = True  # more than 8 candidates on a single GPU: klerg_eval_costs_batch instead of launches of 8

    # shared-psi path: psi once per state-sample pair, the sum over the samples as a tensor-core contraction
    targets_path = "auto"  # "auto" | "fused" | "tensor"
    TENSOR_MIN_TARGETS = 4  # below this the per-target pair pass inside the fused launch is at least as fast

    def _tensor_targets_ok(self, K):
        return self.H <= 64 and K >= self.TENSOR_MIN_TARGETS and not isinstance(self.group, engine.EmulatedShardGroup)

    def _gradient_targets_tensor(self, u):
        """rollout -> forward pass (q_base + q_iter, totals) -> klerg_kl_gradient_targets -> one adjoint launch for all
        targets.  Single rank, H <= 64, K*(D+1) <= 128 per launch (larger K is split)."""
        if self.H > 64:
            raise RuntimeError("the tensor-core targets path needs H <= 64")
        K = self.P.shape[0]
        kmax = min(32, 128 // (self.spec.D + 1))
        ro = engine.rollout(self.dyn, self.bar, self.x0, u, R0=self.R0, want_lin=True)
        traj = ro["traj"][0]
        pre = traj[: self.H]
        v, totals = engine.footprint(self.spec, 0, pre, self.packed, self.n, add_in=self.q_base)
        # sharded: every rank contracts over its own samples; the {sum, max} of q and the per-target gradient partials
        # are all-gathered in rank order (one NCCL call each), so every rank adds the same numbers in the same order
        totals_w = self.group.gather_blocks(totals.reshape(-1)[:2]).reshape(self.group.world, 2).contiguous()
        Pl = ro["P"][0] if ro["P"] is not None else None
        outs = []
        for k0 in range(0, K, kmax):
            gp, _ = engine.kl_gradient_targets(self.spec, pre, self.packed, self.n, v[0], totals_w,
                                               self.P[k0: k0 + kmax], self.floor, want_kl=False)
            gp_w = self.group.gather_blocks(gp).transpose(0, 1).contiguous()  # [K, world, H, D]
            outs.append(engine.adjoint_targets(self.dyn, self.spec, gp_w, ro["dbarr"][0], Pl, traj, u, self.rinv,
                                               self.alpha, self.ctrl_lo, self.ctrl_hi))
        self.evals["grad"] += K
        self.evals["fwd_pairs"] += self.H * self.n
        self.evals["grad_pairs"] += self.H * self.n  # psi is evaluated once for all targets
        dgdx, du, dj, ustar = (torch.cat([o[i] for o in outs]) for i in range(4))
        return dict(du=du, djdlam=dj, u_star=ustar, dgdx=dgdx)

    def q_from(self, v, totals_w):
        """renormalize(q_base + q_iter) given the forward output (for plot_data)."""
        return engine.renormalize_sharded(v[: self.n], totals_w.reshape(totals_w.shape[0], -1)[:, :2], self.floor)

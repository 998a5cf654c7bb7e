"""Device-side operations of the KL-ergodic path on top of the C ABI.

Every function takes/returns CUDA tensors and enqueues work on torch's current
stream; nothing here does arithmetic on the host.  ``ShardGroup`` carries the
(optional) sample-sharding communicator: per-rank partial totals are exchanged
with one small all-gather per phase and combined in rank order on the device.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _cabi as cabi

FLOOR = 1e-6  # renormalize() clamp (klerg_utils.py:45)


def _dev():
    cabi.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


class Workspace:
    """Scratch memory handed to the C side (zero-filled once, reused per stream)."""

    def __init__(self):
        self.G = 0
        self.buf = None

    def get(self, G=1):
        if self.buf is None or G > self.G:
            G = max(int(G), 8, self.G)
            lib = cabi.load()
            nbytes = lib.klerg_workspace_bytes(G)
            old = self.buf
            self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=_dev())
            if old is not None:
                # the head of the workspace carries state that must survive a regrowth: the sticky fault word and the
                # launch / exchange counters of the single-GPU mailbox (the segment records behind it are scratch)
                head = lib.klerg_workspace_bytes(0)
                self.buf[:head].copy_(old[:head])
            self.G = G
        return C.c_void_p(self.buf.data_ptr())


_workspaces = {}


def workspace(G=1):
    key = (torch.cuda.current_device(), cabi.raw_stream())
    ws = _workspaces.get(key)
    if ws is None:
        ws = _workspaces[key] = Workspace()
    return ws.get(G)


class ShardGroup:
    """Sample-sharding communicator (world 1 = no communication)."""

    def __init__(self, process_group=None):
        self.pg = process_group
        if process_group is None:
            self.world, self.rank = 1, 0
        else:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)

    def gather_blocks(self, block):
        """[...] per-rank block -> [world, ...] (rank order), one all-gather."""
        if self.world == 1:
            return block.unsqueeze(0)
        import torch.distributed as dist
        block = block.contiguous()
        flat = torch.empty((self.world * block.shape[0],) + tuple(block.shape[1:]), dtype=block.dtype,
                           device=block.device)  # concatenation along dim 0 (the form every backend accepts)
        dist.all_gather_into_tensor(flat, block, group=self.pg)
        return flat.view((self.world,) + tuple(block.shape))

    def peers(self):
        """klerg_peers for the fused evals: every rank's NVLink mailbox mapped into this process.

        Built once per group: each rank allocates its mailbox (klerg_mailbox_create), the 64-byte
        CUDA-IPC handles travel through the process group's object all-gather, and every rank maps
        the others (klerg_mailbox_open).  Returns None for a single rank."""
        if self.world == 1:
            return None
        if getattr(self, "_peers", None) is not None:
            return self._peers
        cached = _peer_cache.get(id(self.pg))
        if cached is not None and cached[0] is self.pg:  # one mailbox set per process group, however many controllers
            self._peers = cached[1]
            return self._peers
        if self.world > 8:
            raise ValueError(f"the fused evals exchange through at most 8 mailboxes (one box); got world={self.world}")
        import torch.distributed as dist
        lib = cabi.load()
        mine, handle = C.c_void_p(), C.create_string_buffer(64)
        cabi.check(lib.klerg_mailbox_create(C.byref(mine), handle), "klerg_mailbox_create")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=self.pg)
        peers = cabi.Peers()
        peers.world, peers.rank = self.world, self.rank
        for r in range(self.world):
            if r == self.rank:
                peers.mailbox[r] = mine.value
            else:
                ptr = C.c_void_p()
                cabi.check(lib.klerg_mailbox_open(handles[r], C.byref(ptr)), "klerg_mailbox_open")
                peers.mailbox[r] = ptr.value
        torch.cuda.synchronize()
        dist.barrier(group=self.pg)
        self._peers = peers
        _peer_cache[id(self.pg)] = (self.pg, peers)
        return peers

    def max_shard(self, n_total):
        """Samples of the largest shard (every rank sizes its fused grids from it)."""
        return -(-int(n_total) // self.world)

    def close(self):
        """Unmap the peers' mailboxes and free this rank's (collective: call on every rank, after a sync)."""
        cached = _peer_cache.pop(id(self.pg), None)
        peers = cached[1] if cached is not None else getattr(self, "_peers", None)
        self._peers = None
        if peers is None:
            return
        lib = cabi.load()
        torch.cuda.synchronize()
        for r in range(peers.world):
            if peers.mailbox[r]:
                cabi.check(lib.klerg_mailbox_close(C.c_void_p(peers.mailbox[r]), int(r == peers.rank)), "klerg_mailbox_close")
                peers.mailbox[r] = None

    def shard_bounds(self, n_total):
        """Contiguous [lo, hi) slice of the sample axis owned by this rank."""
        base, rem = divmod(int(n_total), self.world)
        lo = self.rank * base + min(self.rank, rem)
        return lo, lo + base + (1 if self.rank < rem else 0)


class EmulatedShardGroup(ShardGroup):
    """One of two ranks of a sample-sharded workspace that both live on ONE GPU (tests of the exchange protocol on a
    single-GPU box): mailboxes are plain device buffers, and the two ranks' fused evals are recorded and launched as
    one cooperative grid by ``emulated_pair`` (separately launched kernels that wait on one another are not
    guaranteed to be co-resident)."""

    def __init__(self, rank, mailboxes):
        self.pg = None
        self.world, self.rank = 2, int(rank)
        self._mail = mailboxes  # keeps the buffers alive
        peers = cabi.Peers()
        peers.world, peers.rank = 2, int(rank)
        for r in range(2):
            peers.mailbox[r] = mailboxes[r].data_ptr()
        self._peers = peers

    def gather_blocks(self, block):
        raise RuntimeError("emulated ranks exchange inside the fused kernels only")

    @staticmethod
    def make_pair(device=None):
        n = cabi.load().klerg_mailbox_bytes()
        mail = [torch.zeros(n, dtype=torch.uint8, device=device or _dev()) for _ in range(2)]
        return EmulatedShardGroup(0, mail), EmulatedShardGroup(1, mail)


def emulated_pair(call_rank0, call_rank1, grid_limit=64):
    """Run ONE fused eval of each emulated rank (the two callables issue exactly one klerg_eval_* call each) as a
    single cooperative launch; returns the callables' results."""
    lib = cabi.load()
    cabi.check(lib.klerg_set_option(cabi.OPT_GRID_LIMIT, int(grid_limit)), "klerg_set_option")
    try:
        cabi.check(lib.klerg_emu_begin(), "klerg_emu_begin")
        out = (call_rank0(), call_rank1())
        cabi.check(lib.klerg_emu_launch(cabi.stream_ptr()), "klerg_emu_launch")
    finally:
        lib.klerg_set_option(cabi.OPT_GRID_LIMIT, 0)
    return out


_peer_cache = {}
SINGLE = ShardGroup()


class nvtx_range:
    """NVTX range around a phase of Robot.step() (visible in nsys / ncu --nvtx timelines).  KLERG_NVTX=0 turns them off."""
    enabled = os.environ.get("KLERG_NVTX", "1") != "0"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if nvtx_range.enabled:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if nvtx_range.enabled:
            torch.cuda.nvtx.range_pop()
        return False


def set_eval_overlap(on):
    """Declare consecutive fused evals on a stream independent (no eval reads what the previous one wrote): the
    next eval's CTAs then start while the previous eval's last CTA is still in its gather + adjoint tail.  Batches
    of independent evals (bench, candidate sweeps) only; the planner's evals are separated by host decisions."""
    cabi.check(cabi.load().klerg_set_option(cabi.OPT_EVAL_OVERLAP, int(bool(on))), "klerg_set_option")


_saturate_milli = [0]


def set_saturate(app_thresh):
    """u* of the evals enqueued from now on: 0 = clamp to the control limits (klerg.py:522), t > 0 = the reference's
    ``saturate`` flag, tanh(u / t) * control_lim[:, 1] (Robot.saturate_control, klerg.py:342-349)."""
    milli = int(round(1000.0 * float(app_thresh)))
    if milli != _saturate_milli[0]:
        cabi.check(cabi.load().klerg_set_option(cabi.OPT_SATURATE_MILLI, milli), "klerg_set_option")
        _saturate_milli[0] = milli


def padded(n):
    return max(4, (int(n) + 3) // 4 * 4)


def pack_samples(spec, samples):
    """[N,D] raw samples -> packed/scaled SoA [D, ld] (klerg_pack_samples)."""
    n = samples.shape[0]
    ld = padded(n)
    packed = torch.empty((spec.D, ld), dtype=torch.float32, device=samples.device)
    cabi.check(cabi.load().klerg_pack_samples(C.byref(spec), cabi.ptr(samples), n, cabi.ptr(packed), ld,
                                              cabi.stream_ptr()), "klerg_pack_samples")
    return packed


def footprint(spec, mode, states, packed, n, add_in=None, out=None):
    """states [G,T,S] (or [T,S]) -> (out [G, ld], totals [G,2] float64)."""
    if states.dim() == 2:
        states = states.unsqueeze(0)
    G, T, S = states.shape
    assert S == spec.S, (S, spec.S)
    if T > 0 and not (states.stride(2) == 1 and states.stride(1) == S):
        states = states.contiguous()  # rows must be dense; segments may be strided views
    seg_stride = states.stride(0) if (T > 0 and G > 1) else T * S
    states_ptr = C.c_void_p(states.data_ptr()) if T > 0 else None
    ld = packed.shape[1]
    if out is None:
        out = torch.empty((G, ld), dtype=torch.float32, device=packed.device)
    totals = torch.empty((G, 2), dtype=torch.float64, device=packed.device)
    cabi.check(cabi.load().klerg_footprint(
        C.byref(spec), int(mode), states_ptr, G, T, seg_stride, cabi.ptr(packed), int(n), ld,
        cabi.ptr(add_in), cabi.ptr(out), out.stride(0), cabi.ptr(totals), workspace(G), cabi.stream_ptr()),
        "klerg_footprint")
    return out, totals


FOOTPRINT_TC = os.environ.get("KLERG_FOOTPRINT_TC", "1") != "0"  # the tensor-core form of the history / spread pass
FOOTPRINT_TC_MIN_PAIRS = 1 << 28  # below this the CUDA-core pass is at least as fast (launches, chunk packing)
_tc_scratch = {}


def footprint_sum_max(spec, states, t_sum, packed, n, tensor_cores=None):
    """ONE pass over the squared distances for both memory-buffer passes of a planner step: states [T,S] with the drawn
    rows first -> (sum over the first t_sum rows [ld], max over all rows [ld], totals [2] of the sum).  Large passes
    (history of a 1e7-sample workspace) take the tensor-core form (klerg_footprint_sum_max_tc: distances as one K = 8
    tf32 MMA step, min / exp / add on the CUDA cores), which falls back on the device when the states leave the radius
    of the expanded pair form; ``tensor_cores`` forces the choice (tests)."""
    states = states.contiguous()
    T, S = states.shape
    assert S == spec.S, (S, spec.S)
    ld = packed.shape[1]
    out_sum = torch.empty(ld, dtype=torch.float32, device=packed.device)
    out_max = torch.empty(ld, dtype=torch.float32, device=packed.device)
    totals = torch.empty(2, dtype=torch.float64, device=packed.device)
    # long state lists only: a 128-sample tile pays its fixed work (A rows into TMEM, two CTA-wide meetings) once per
    # T / 128 chunks
    use_tc = tensor_cores if tensor_cores is not None else (FOOTPRINT_TC and spec.D <= 6 and T >= 1024
                                                           and T * int(n) >= FOOTPRINT_TC_MIN_PAIRS)
    if use_tc and T > 0 and n > 0:
        lib = cabi.load()
        need = int(lib.klerg_footprint_tc_scratch_bytes(T))
        key = (packed.device.index, cabi.raw_stream())
        scratch = _tc_scratch.get(key)
        if scratch is None or scratch.numel() < need:
            scratch = _tc_scratch[key] = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=packed.device)
        cabi.check(lib.klerg_footprint_sum_max_tc(
            C.byref(spec), cabi.ptr(states), T, int(t_sum), cabi.ptr(packed), int(n), ld, cabi.ptr(out_sum),
            cabi.ptr(out_max), cabi.ptr(totals), workspace(1), cabi.ptr(scratch), scratch.numel(), cabi.stream_ptr()),
            "klerg_footprint_sum_max_tc")
        return out_sum, out_max, totals
    cabi.check(cabi.load().klerg_footprint_sum_max(
        C.byref(spec), cabi.ptr(states) if T > 0 else None, T, int(t_sum), cabi.ptr(packed), int(n), ld, cabi.ptr(out_sum),
        cabi.ptr(out_max), cabi.ptr(totals), workspace(1), cabi.stream_ptr()), "klerg_footprint_sum_max")
    return out_sum, out_max, totals


def vector_stats(x):
    stats = torch.empty(4, dtype=torch.float64, device=x.device)
    cabi.check(cabi.load().klerg_vector_stats(cabi.ptr(x), x.numel(), cabi.ptr(stats), workspace(), cabi.stream_ptr()),
               "klerg_vector_stats")
    return stats


def renormalize(x, floor=FLOOR):
    out = torch.empty_like(x)
    cabi.check(cabi.load().klerg_renormalize(cabi.ptr(x), x.numel(), float(floor), cabi.ptr(out), workspace(),
                                             cabi.stream_ptr()), "klerg_renormalize")
    return out


def renormalize_sharded(x, totals_w, floor=FLOOR):
    """renormalize() of a sample-sharded vector: totals_w [world,2] = per-rank {sum, max}."""
    stats = combine_blocks(totals_w.contiguous(), [cabi.RED_SUM, cabi.RED_MAX])
    x = x.contiguous()
    out = torch.empty_like(x)
    cabi.check(cabi.load().klerg_renormalize_with_stats(cabi.ptr(x), x.numel(), cabi.ptr(stats), float(floor),
                                                        cabi.ptr(out), cabi.stream_ptr()), "klerg_renormalize_with_stats")
    return out


def cost_norm_(x):
    cabi.check(cabi.load().klerg_cost_norm(cabi.ptr(x), x.numel(), workspace(), cabi.stream_ptr()), "klerg_cost_norm")
    return x


def kl_gradient(spec, states, packed, n, w):
    states = states.contiguous()
    H = states.shape[0]
    dgdx = torch.empty((H, spec.S), dtype=torch.float32, device=packed.device)
    cabi.check(cabi.load().klerg_kl_gradient(C.byref(spec), cabi.ptr(states), H, cabi.ptr(packed), int(n),
                                             packed.shape[1], cabi.ptr(w), cabi.ptr(dgdx), workspace(),
                                             cabi.stream_ptr()), "klerg_kl_gradient")
    return dgdx


def kl_gradient_fused(spec, states, packed, n, v, totals_w, p, floor=FLOOR):
    """-> (grad_part [H,D] float64, kl_part [2] float64) for this rank's samples."""
    states = states.contiguous()
    H = states.shape[0]
    grad_part = torch.empty((H, spec.D), dtype=torch.float64, device=packed.device)
    kl_part = torch.empty(2, dtype=torch.float64, device=packed.device)
    cabi.check(cabi.load().klerg_kl_gradient_fused(
        C.byref(spec), cabi.ptr(states), H, cabi.ptr(packed), int(n), packed.shape[1], cabi.ptr(v),
        cabi.ptr(totals_w), totals_w.shape[0], cabi.ptr(p), float(floor), cabi.ptr(grad_part), cabi.ptr(kl_part),
        workspace(), cabi.stream_ptr()), "klerg_kl_gradient_fused")
    return grad_part, kl_part


_targets_scratch = {}


def kl_gradient_targets(spec, states, packed, n, v, totals_w, P, floor=FLOOR, want_kl=True):
    """K belief targets, psi shared (tensor-core contraction): -> (grad_parts [K,H,D] float64, kl_parts [K,2] float64
    or None) for this rank's samples.  P [K, ld] rows padded to the sample stride.  ``want_kl=False`` skips the
    per-target KL terms (a log per target and sample)."""
    lib = cabi.load()
    states = states.contiguous()
    H, K = states.shape[0], P.shape[0]
    dev = packed.device
    nbytes = lib.klerg_kl_gradient_targets_scratch_bytes(H, K)
    key = (dev.index, cabi.raw_stream())
    scr = _targets_scratch.get(key)
    if scr is None or scr[0].numel() < nbytes:
        scr = _targets_scratch[key] = (torch.empty(nbytes, dtype=torch.uint8, device=dev),
                                       torch.zeros(1, dtype=torch.int32, device=dev))
    grad_parts = torch.empty((K, H, spec.D), dtype=torch.float64, device=dev)
    kl_parts = torch.empty((K, 2), dtype=torch.float64, device=dev) if want_kl else None
    cabi.check(lib.klerg_kl_gradient_targets(
        C.byref(spec), cabi.ptr(states), H, cabi.ptr(packed), int(n), packed.shape[1], cabi.ptr(v), cabi.ptr(totals_w),
        totals_w.shape[0], cabi.ptr(P), K, P.shape[1], float(floor), cabi.ptr(grad_parts), cabi.ptr(kl_parts),
        cabi.ptr(scr[0]), cabi.ptr(scr[1]), cabi.stream_ptr()), "klerg_kl_gradient_targets")
    return grad_parts, kl_parts


def targets_gradient_fault():
    """True if any klerg_kl_gradient_targets launch reported a timed-out in-kernel wait (one small D2H read each)."""
    return any(int(f.item()) != 0 for _, f in _targets_scratch.values())


def kl_cost(v, n, totals_w, p, p_stats, barrier_sum, group=SINGLE, floor=FLOOR):
    """KL(p||q) + barrier for G candidates: v [G, ld] -> cost [G]."""
    G = v.shape[0]
    kl_part = torch.empty((G, 2), dtype=torch.float64, device=v.device)
    lib = cabi.load()
    cabi.check(lib.klerg_kl_cost_partial(cabi.ptr(v), v.stride(0), G, int(n), cabi.ptr(totals_w), totals_w.shape[0],
                                         cabi.ptr(p), float(floor), cabi.ptr(kl_part), workspace(G),
                                         cabi.stream_ptr()), "klerg_kl_cost_partial")
    parts = group.gather_blocks(kl_part)
    cost = torch.empty(G, dtype=torch.float32, device=v.device)
    cabi.check(lib.klerg_kl_cost_final(cabi.ptr(parts), parts.shape[0], G, cabi.ptr(p_stats), cabi.ptr(barrier_sum),
                                       cabi.ptr(cost), cabi.stream_ptr()), "klerg_kl_cost_final")
    return cost


def combine_blocks(blocks, kinds):
    """[world, n] float64 -> [n] with per-column sum/max/min (rank order)."""
    world, n = blocks.shape
    out = torch.empty(n, dtype=torch.float64, device=blocks.device)
    arr = (C.c_int * n)(*kinds)
    cabi.check(cabi.load().klerg_combine_blocks(cabi.ptr(blocks.contiguous()), world, n, arr, cabi.ptr(out),
                                                cabi.stream_ptr()), "klerg_combine_blocks")
    return out


def target_weight(mode, samples, lim_lo, lim_hi, spread, p_raw, n_total, temp, renorm, group=SINGLE, floor=FLOOR):
    """get_target_dist weighting (klerg.py:452-486) -> (p [N], p_stats [1] float64).

    mode 0: p ** mean(spread')   mode 1: p + (1-spread')*min p   mode 2: unchanged.
    ``spread`` None = empty memory buffer.  ``renorm`` applies renormalize().
    """
    lib = cabi.load()
    n, D = samples.shape
    lo, hi = cabi.farr(lim_lo), cabi.farr(lim_hi)
    dev = samples.device
    acc1 = torch.empty(4, dtype=torch.float64, device=dev)
    cabi.check(lib.klerg_target_stage1(cabi.ptr(samples), D, n, lo, hi, cabi.ptr(spread), cabi.ptr(p_raw),
                                       cabi.ptr(acc1), workspace(), cabi.stream_ptr()), "klerg_target_stage1")
    acc1 = combine_blocks(group.gather_blocks(acc1), [cabi.RED_MAX, cabi.RED_SUM, cabi.RED_SUM, cabi.RED_MIN])
    expo = torch.empty(3, dtype=torch.float64, device=dev)
    cabi.check(lib.klerg_target_exponent(cabi.ptr(acc1), int(n_total), cabi.ptr(expo), cabi.stream_ptr()),
               "klerg_target_exponent")
    p2 = torch.empty(n, dtype=torch.float32, device=dev)
    acc2 = torch.empty(2, dtype=torch.float64, device=dev)
    cabi.check(lib.klerg_target_stage2(int(mode), cabi.ptr(samples), D, n, lo, hi, cabi.ptr(spread), cabi.ptr(p_raw),
                                       cabi.ptr(expo), cabi.ptr(p2), cabi.ptr(acc2), workspace(), cabi.stream_ptr()),
               "klerg_target_stage2")
    acc2 = combine_blocks(group.gather_blocks(acc2), [cabi.RED_SUM, cabi.RED_MAX])
    p = torch.empty(n, dtype=torch.float32, device=dev)
    p_stats = torch.empty(1, dtype=torch.float64, device=dev)
    cabi.check(lib.klerg_target_stage3(cabi.ptr(p2), n, cabi.ptr(acc2), int(bool(renorm)), float(floor), float(temp),
                                       cabi.ptr(p), cabi.ptr(p_stats), workspace(), cabi.stream_ptr()),
               "klerg_target_stage3")
    p_stats = combine_blocks(group.gather_blocks(p_stats), [cabi.RED_SUM])
    return p, p_stats, expo


def rollout(dyn, bar, x0, u, R0=None, want_lin=False, want_R=False):
    """u [B,H,A] (or [H,A]) -> dict(traj [B,H+1,S], barrier [B], dbarr, P, R)."""
    if u.dim() == 2:
        u = u.unsqueeze(0)
    u = u.contiguous()
    B, H, A = u.shape
    dev = u.device
    S = dyn.S
    traj = torch.empty((B, H + 1, S), dtype=torch.float32, device=dev)
    bsum = torch.empty(B, dtype=torch.float32, device=dev)
    dbarr = torch.empty((B, H, S), dtype=torch.float32, device=dev) if want_lin else None
    P = torch.empty((B, H, A * A), dtype=torch.float32, device=dev) if (want_lin and dyn.kind == cabi.DYN_ROLL) else None
    R_out = torch.empty((B, 9), dtype=torch.float32, device=dev) if want_R else None
    cabi.check(cabi.load().klerg_rollout(C.byref(dyn), C.byref(bar) if bar is not None else None, cabi.ptr(x0),
                                         cabi.ptr(R0), cabi.ptr(u), B, H, cabi.ptr(traj), cabi.ptr(bsum),
                                         cabi.ptr(dbarr), cabi.ptr(P), cabi.ptr(R_out), cabi.stream_ptr()),
               "klerg_rollout")
    return dict(traj=traj, barrier=bsum, dbarr=dbarr, P=P, R=R_out)


def barrier_eval(bar, x, want_value=True, want_grad=True):
    x = x.contiguous()
    T, S = x.shape
    value = torch.empty(T, dtype=torch.float32, device=x.device) if want_value else None
    grad = torch.empty((T, S), dtype=torch.float32, device=x.device) if want_grad else None
    cabi.check(cabi.load().klerg_barrier_eval(C.byref(bar) if bar is not None else None, cabi.ptr(x), T, S,
                                              cabi.ptr(value), cabi.ptr(grad), cabi.stream_ptr()), "klerg_barrier_eval")
    return value, grad


def barrier_eval_ext(bar, x, x_ref=None, tilt=None, want_value=True, want_grad=True):
    """BarrierFunction with limits relative to ``x_ref`` rows (VelocityBarrier) and / or the tilt-dependent yaw limits
    and tilt term of TiltBarrierFunction -> (value [T], grad [T,S], tilt [T] or None)."""
    x = x.contiguous()
    T, S = x.shape
    value = torch.empty(T, dtype=torch.float32, device=x.device) if want_value else None
    grad = torch.empty((T, S), dtype=torch.float32, device=x.device) if want_grad else None
    tilt_out = torch.empty(T, dtype=torch.float32, device=x.device) if tilt is not None else None
    cabi.check(cabi.load().klerg_barrier_eval_ext(
        C.byref(bar) if bar is not None else None, cabi.ptr(x), cabi.ptr(x_ref.contiguous()) if x_ref is not None else None,
        C.byref(tilt) if tilt is not None else None, T, S, cabi.ptr(value), cabi.ptr(grad), cabi.ptr(tilt_out),
        cabi.stream_ptr()), "klerg_barrier_eval_ext")
    return value, grad, tilt_out


def adjoint(dyn, spec, grad_parts, dbarr, P, traj, u, rinv, alpha, ctrl_lo, ctrl_hi):
    """grad_parts [world,H,D] float64 -> dgdx [H,S], du [H,A], djdlam [H], u_star [H,A]."""
    world, H, D = grad_parts.shape
    dev = grad_parts.device
    S, A = dyn.S, dyn.A
    dgdx = torch.empty((H, S), dtype=torch.float32, device=dev)
    du = torch.empty((H, A), dtype=torch.float32, device=dev)
    dj = torch.empty(H, dtype=torch.float32, device=dev)
    ustar = torch.empty((H, A), dtype=torch.float32, device=dev)
    cabi.check(cabi.load().klerg_adjoint(
        C.byref(dyn), C.byref(spec), H, cabi.ptr(grad_parts.contiguous()), world, cabi.ptr(dbarr.contiguous()),
        cabi.ptr(P), cabi.ptr(traj.contiguous()), cabi.ptr(u.contiguous()), cabi.farr(rinv), float(alpha),
        cabi.farr(ctrl_lo), cabi.farr(ctrl_hi), cabi.ptr(dgdx), cabi.ptr(du), cabi.ptr(dj), cabi.ptr(ustar),
        cabi.stream_ptr()), "klerg_adjoint")
    return dgdx, du, dj, ustar


def policy_rollout(dyn, pol, x0, R0, u_in, H):
    """Closed-loop rollout under a state-feedback default policy (klerg.py:409-431 with default_policies.py:53-119):
    -> (u_eff [H,A] the controls the policy applied, dmudx [H,A,S])."""
    dev = x0.device
    u_eff = torch.empty((H, dyn.A), dtype=torch.float32, device=dev)
    dmudx = torch.empty((H, dyn.A, dyn.S), dtype=torch.float32, device=dev)
    cabi.check(cabi.load().klerg_policy_rollout(C.byref(dyn), C.byref(pol), cabi.ptr(x0), cabi.ptr(R0),
                                                cabi.ptr(u_in.contiguous()) if u_in is not None else None, int(H),
                                                cabi.ptr(u_eff), cabi.ptr(dmudx), cabi.stream_ptr()), "klerg_policy_rollout")
    return u_eff, dmudx


def adjoint_policy(dyn, dgdx, dbarr, P, dmudx, u, rinv, alpha, ctrl_lo, ctrl_hi):
    """Adjoint sweep with the closed-loop linearisation A_t + B dmudx_t -> du [H,A], djdlam [H], u_star [H,A]."""
    H = dgdx.shape[0]
    dev = dgdx.device
    du = torch.empty((H, dyn.A), dtype=torch.float32, device=dev)
    dj = torch.empty(H, dtype=torch.float32, device=dev)
    ustar = torch.empty((H, dyn.A), dtype=torch.float32, device=dev)
    cabi.check(cabi.load().klerg_adjoint_policy(
        C.byref(dyn), H, cabi.ptr(dgdx.contiguous()), cabi.ptr(dbarr.contiguous()), cabi.ptr(P), cabi.ptr(dmudx.contiguous()),
        cabi.ptr(u.contiguous()), cabi.farr(rinv), float(alpha), cabi.farr(ctrl_lo), cabi.farr(ctrl_hi), cabi.ptr(du),
        cabi.ptr(dj), cabi.ptr(ustar), cabi.stream_ptr()), "klerg_adjoint_policy")
    return du, dj, ustar


def adjoint_targets(dyn, spec, grad_parts, dbarr, P, traj, u, rinv, alpha, ctrl_lo, ctrl_hi):
    """grad_parts [K,world,H,D] float64 -> dgdx [K,H,S], du [K,H,A], djdlam [K,H], u_star [K,H,A] (one launch)."""
    K, world, H, D = grad_parts.shape
    dev = grad_parts.device
    S, A = dyn.S, dyn.A
    dgdx = torch.empty((K, H, S), dtype=torch.float32, device=dev)
    du = torch.empty((K, H, A), dtype=torch.float32, device=dev)
    dj = torch.empty((K, H), dtype=torch.float32, device=dev)
    ustar = torch.empty((K, H, A), dtype=torch.float32, device=dev)
    cabi.check(cabi.load().klerg_adjoint_targets(
        C.byref(dyn), C.byref(spec), H, K, cabi.ptr(grad_parts.contiguous()), world, cabi.ptr(dbarr.contiguous()),
        cabi.ptr(P), cabi.ptr(traj.contiguous()), cabi.ptr(u.contiguous()), cabi.farr(rinv), float(alpha),
        cabi.farr(ctrl_lo), cabi.farr(ctrl_hi), cabi.ptr(dgdx), cabi.ptr(du), cabi.ptr(dj), cabi.ptr(ustar),
        cabi.stream_ptr()), "klerg_adjoint_targets")
    return dgdx, du, dj, ustar


class EvalBuffers:
    """Caller-owned outputs/scratch of the fused evals for one (H, S, A, ld) shape.

    Two alternating output sets, so the results of the previous gradient eval stay valid
    while the next one runs (the planner keeps the last accepted iterate for plot_data)."""

    def __init__(self, H, S, A, ld, device, n_sets=2, max_g=8):
        f32 = dict(dtype=torch.float32, device=device)
        f64 = dict(dtype=torch.float64, device=device)
        self.H, self.S, self.A, self.ld, self.max_g = H, S, A, ld, max_g
        self.sets = []
        for _ in range(n_sets):
            # djdlam and u_star are what the host control flow reads after every gradient eval: one buffer, one D2H;
            # the last float is the kernel's copy of the sticky fault word (a timed-out in-kernel wait)
            host_pack = torch.zeros(H + H * A + 1, **f32)
            self.sets.append(dict(
                v=torch.empty(ld, **f32), traj=torch.empty((H + 1, S), **f32), totals=torch.empty((1, 2), **f64),
                cost=torch.empty(1, **f32), dgdx=torch.empty((H, S), **f32), du=torch.empty((H, A), **f32),
                host_pack=host_pack, djdlam=host_pack[:H], u_star=host_pack[H:H + H * A].view(H, A),
                fault=host_pack[H + H * A:], kl=torch.empty(2, **f64)))
        self.turn = 0
        self.v_costs = torch.empty((max_g, ld), **f32)
        self.cost_pack = torch.zeros(max_g + 1, **f32)  # costs of one launch + fault word: one D2H

    def next_set(self):
        self.turn = (self.turn + 1) % len(self.sets)
        return self.sets[self.turn]


def eval_gradient(spec, dyn, bar, peers, x0, R0, u, packed, n, q_base, p, p_stats, rinv, alpha, ctrl_lo, ctrl_hi, out,
                  floor=FLOOR, want_cost=False):
    """One fused launch: rollout + footprint + renormalize + gradient + adjoint (klerg_eval_gradient).

    ``want_cost`` also accumulates the KL of the differentiated footprint (not needed by the planner)."""
    H = u.shape[-2]
    cabi.check(cabi.load().klerg_eval_gradient(
        C.byref(spec), C.byref(dyn), C.byref(bar) if bar is not None else None, peers, cabi.ptr(x0), cabi.ptr(R0),
        cabi.ptr(u), H, cabi.ptr(packed), int(n), packed.shape[1], cabi.ptr(q_base), cabi.ptr(p), cabi.ptr(p_stats),
        float(floor), rinv, float(alpha), ctrl_lo, ctrl_hi, cabi.ptr(out["v"]), cabi.ptr(out["traj"]),
        cabi.ptr(out["totals"]), cabi.ptr(out["cost"]) if want_cost else None, cabi.ptr(out["dgdx"]),
        cabi.ptr(out["du"]), cabi.ptr(out["djdlam"]), cabi.ptr(out["u_star"]),
        cabi.ptr(out["kl"]) if want_cost else None, cabi.ptr(out.get("fault")), workspace(8), cabi.stream_ptr()),
        "klerg_eval_gradient")
    return out


def eval_gradient_targets(spec, dyn, bar, peers, x0, R0, u, packed, n, q_base, P, P_stats, rinv, alpha, ctrl_lo, ctrl_hi,
                          v_scratch, floor=FLOOR):
    """One fused launch for K belief targets P [K, stride]: shared rollout / forward pass / q, per-target gradient
    and adjoint (klerg_eval_gradient_targets) -> dict of [K, ...] tensors."""
    K, stride = P.shape
    H, A, S = u.shape[-2], dyn.A, dyn.S
    dev = P.device
    f32 = dict(dtype=torch.float32, device=dev)
    out = dict(dgdx=torch.empty((K, H, S), **f32), du=torch.empty((K, H, A), **f32), djdlam=torch.empty((K, H), **f32),
               u_star=torch.empty((K, H, A), **f32), traj=torch.empty((H + 1, S), **f32),
               totals=torch.empty((1, 2), dtype=torch.float64, device=dev))
    cabi.check(cabi.load().klerg_eval_gradient_targets(
        C.byref(spec), C.byref(dyn), C.byref(bar) if bar is not None else None, peers, cabi.ptr(x0), cabi.ptr(R0),
        cabi.ptr(u), H, cabi.ptr(packed), int(n), packed.shape[1], cabi.ptr(q_base), cabi.ptr(P), K, stride,
        cabi.ptr(P_stats), float(floor), rinv, float(alpha), ctrl_lo, ctrl_hi, cabi.ptr(v_scratch), cabi.ptr(out["traj"]),
        cabi.ptr(out["totals"]), None, cabi.ptr(out["dgdx"]), cabi.ptr(out["du"]), cabi.ptr(out["djdlam"]),
        cabi.ptr(out["u_star"]), None, None, workspace(8), cabi.stream_ptr()), "klerg_eval_gradient_targets")
    return out


def eval_costs(spec, dyn, bar, peers, x0, R0, U, packed, n, q_base, p, p_stats, v_scratch, cost, floor=FLOOR, fault=None):
    """One fused launch: get_cost of G <= 8 candidates U [G,H,A] -> cost [G] (klerg_eval_costs)."""
    G, H, _ = U.shape
    cabi.check(cabi.load().klerg_eval_costs(
        C.byref(spec), C.byref(dyn), C.byref(bar) if bar is not None else None, peers, cabi.ptr(x0), cabi.ptr(R0),
        cabi.ptr(U), G, H, cabi.ptr(packed), int(n), packed.shape[1], cabi.ptr(q_base), cabi.ptr(p), cabi.ptr(p_stats),
        float(floor), cabi.ptr(v_scratch), None, None, cabi.ptr(cost), cabi.ptr(fault), workspace(8), cabi.stream_ptr()),
        "klerg_eval_costs")
    return cost


def plan_optimize(spec, dyn, bar, peers, x0, R0, u, packed, n, q_base, p, p_stats, rinv, alpha, ctrl_lo, ctrl_hi, num_iters,
                  fixed_lam, lam, buf, floor=FLOOR, max_app_dur=5):
    """The planner's optimisation loop enqueued at once and decided on the device (klerg_plan_optimize).  ``u`` [H,A]
    (device) is updated in place; returns the packed result on the device (one D2H for the caller):
    {last_cost, fault, cost evals, gradient evals, accepted iterations, 0, 0, 0, u[H*A], traj[(H+1)*S]}."""
    lib = cabi.load()
    H, A, S = u.shape[-2], dyn.A, dyn.S
    ld = packed.shape[1]
    if getattr(buf, "plan_scratch", None) is None:
        buf.plan_scratch = torch.empty(lib.klerg_plan_scratch_bytes(H, S, A, ld), dtype=torch.uint8, device=u.device)
        buf.plan_result = torch.zeros(lib.klerg_plan_result_floats(H, S, A), dtype=torch.float32, device=u.device)
    cabi.check(lib.klerg_plan_optimize(
        C.byref(spec), C.byref(dyn), C.byref(bar) if bar is not None else None, peers, cabi.ptr(x0), cabi.ptr(R0),
        cabi.ptr(u), H, cabi.ptr(packed), int(n), ld, cabi.ptr(q_base), cabi.ptr(p), cabi.ptr(p_stats), float(floor), rinv,
        float(alpha), ctrl_lo, ctrl_hi, int(num_iters), int(bool(fixed_lam)), int(lam), int(max_app_dur),
        cabi.ptr(buf.plan_scratch), cabi.ptr(buf.plan_result), workspace(8), cabi.stream_ptr()), "klerg_plan_optimize")
    return buf.plan_result


def eval_costs_batch(spec, dyn, bar, x0, R0, U, packed, n, q_base, p, p_stats, floor=FLOOR):
    """ONE launch: get_cost of any number of candidates U [B,H,A] -> cost [B] (klerg_eval_costs_batch; single GPU)."""
    lib = cabi.load()
    B, H, _ = U.shape
    dev = U.device
    ld = packed.shape[1]
    v = torch.empty((B, ld), dtype=torch.float32, device=dev)
    scratch = torch.empty(lib.klerg_eval_costs_batch_scratch_bytes(B, H, spec.D), dtype=torch.uint8, device=dev)
    pack = torch.zeros(B + 1, dtype=torch.float32, device=dev)  # costs + fault word
    cabi.check(lib.klerg_eval_costs_batch(
        C.byref(spec), C.byref(dyn), C.byref(bar) if bar is not None else None, cabi.ptr(x0), cabi.ptr(R0), cabi.ptr(U), B, H,
        cabi.ptr(packed), int(n), ld, cabi.ptr(q_base), cabi.ptr(p), cabi.ptr(p_stats), float(floor), cabi.ptr(v),
        cabi.ptr(scratch), None, cabi.ptr(pack), cabi.ptr(pack[B:]), workspace(8), cabi.stream_ptr()), "klerg_eval_costs_batch")
    return pack[:B], pack[B:]


def debug_stamps():
    """SM-cycle stamps of the last fused gradient eval on the current stream's workspace (8 int64)."""
    key = (torch.cuda.current_device(), cabi.raw_stream())
    ws = _workspaces[key]
    off = cabi.load().klerg_debug_stamps_offset()
    return ws.buf[off:off + 144].view(torch.int64).cpu().tolist()


def debug_cta_stamps(peers=None):
    """[160, 8] globaltimer stamps (ns) per CTA of the last fused gradient eval (KLERG_STAMPS builds).  ``peers``: the
    klerg_peers of a sharded context - the stamps then live in this rank's NVLink mailbox, not in the workspace."""
    lib = cabi.load()
    if peers is not None:
        import numpy as np
        from cuda import cudart
        mb_off = lib.klerg_debug_cta_stamps_offset() - (lib.klerg_fused_fault_offset() - 20) - 256  # MB_OFF_DBG
        host = np.zeros(160 * 8, dtype=np.int64)
        torch.cuda.synchronize()
        (err,) = cudart.cudaMemcpy(host.ctypes.data, int(peers.mailbox[peers.rank]) + mb_off, host.nbytes,
                                   cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        if int(err) != 0:
            raise RuntimeError(f"cudaMemcpy of the CTA stamps failed: {err}")
        return torch.from_numpy(host).view(160, 8)
    key = (torch.cuda.current_device(), cabi.raw_stream())
    ws = _workspaces[key]
    off = lib.klerg_debug_cta_stamps_offset()
    return ws.buf[off:off + 160 * 64].view(torch.int64).view(160, 8).cpu()


def fused_fault():
    """True if a fused eval on the current stream's workspace gave up waiting at a meeting point (one small D2H read;
    the planner reads the kernels' own copy of the word with the results it fetches anyway)."""
    key = (torch.cuda.current_device(), cabi.raw_stream())
    ws = _workspaces.get(key)
    if ws is None or ws.buf is None:
        return False
    off = cabi.load().klerg_fused_fault_offset()
    return bool(ws.buf[off:off + 4].view(torch.int32).item())


# torch.get_rng_state() of the CPU generator (ATen CPUGeneratorImplStateLegacy): seed u64, left i32 @8, seeded i32,
# next u64 @16, state u64[624] @24, then the normal-distribution cache
_RNG_LEFT, _RNG_NEXT, _RNG_STATE, _RNG_WORDS = 8, 16, 24, 624


def _rng_fields(raw):
    import numpy as np
    left = int(raw[_RNG_LEFT:_RNG_LEFT + 4].view(np.int32)[0])
    nxt = int(raw[_RNG_NEXT:_RNG_NEXT + 8].view(np.uint64)[0])
    words = torch.from_numpy(raw[_RNG_STATE:_RNG_STATE + 8 * _RNG_WORDS].view(np.uint64).astype(np.uint32).view(np.int32))
    return left, nxt, words


def _rng_set_from(raw, back):
    """Put the state words the kernel handed back (624 words, left, next) into torch's CPU generator."""
    import numpy as np
    new = raw.copy()
    new[_RNG_LEFT:_RNG_LEFT + 4] = np.array([back[_RNG_WORDS]], dtype=np.int32).view(np.uint8)
    new[_RNG_NEXT:_RNG_NEXT + 8] = np.array([back[_RNG_WORDS + 1]], dtype=np.uint64).view(np.uint8)
    new[_RNG_STATE:_RNG_STATE + 8 * _RNG_WORDS] = back[:_RNG_WORDS].astype(np.uint64).view(np.uint8)
    torch.set_rng_state(torch.from_numpy(new))


def _enqueue_uniform(blob, n_rows, low, span, row_lo, row_hi, dev):
    """klerg_mt19937_uniform on the current stream from the generator state ``blob`` -> (rows, advanced state words)."""
    left, nxt, words = _rng_fields(blob.numpy())
    st_in = words.to(dev, non_blocking=True)
    st_out = torch.empty(_RNG_WORDS + 2, dtype=torch.int32, device=dev)
    out = torch.empty((max(row_hi - row_lo, 0), low.numel()), dtype=torch.float32, device=dev)
    cabi.check(cabi.load().klerg_mt19937_uniform(
        cabi.ptr(st_in), left, nxt, int(n_rows), low.numel(), cabi.farr(low.tolist()), cabi.farr(span.tolist()), int(row_lo),
        int(row_hi), cabi.ptr(out) if out.numel() else None, cabi.ptr(st_out), cabi.stream_ptr()), "klerg_mt19937_uniform")
    return out, st_out


def _dev_index(dev):
    dev = torch.device(dev)
    return torch.cuda.current_device() if dev.index is None else dev.index


class UniformPrefetch:
    """The NEXT step's workspace draw, enqueued speculatively on a side stream from the generator state the current
    step leaves behind.  The draw of step k+1 is a pure function of (generator state, row count, box); if the next
    call finds torch's generator exactly where this one expected it (byte comparison of the 5 KB state) and asks for
    the same draw, the rows are already there and the generator jumps to the recorded end state - bit for bit what
    the in-line draw would have produced.  Anything else (another consumer of the generator in between, a different
    sample count or box) discards the speculation.  One serial CTA draws 6e7 numbers in 44 ms: this takes it off the
    critical path of Robot.step() (it overlaps the history pass of the step before)."""

    def __init__(self):
        self.pending = None
        self.stream = None
        self.back = None
        self.hits = self.misses = 0

    def launch(self, n_rows, low, high, row_lo, row_hi, device):
        dev = device or _dev()
        low = torch.as_tensor(low, dtype=torch.float32).reshape(-1)
        span = torch.as_tensor(high, dtype=torch.float32).reshape(-1) - low
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev)
        if self.pending is not None:  # an unconsumed speculation: let it finish before its hand-back buffer is reused
            self.pending["ev"].synchronize()
            self.pending = None
        blob = torch.get_rng_state()
        main = torch.cuda.current_stream(dev)
        self.stream.wait_stream(main)  # not before what is already enqueued (keeps allocator reuse simple)
        with torch.cuda.stream(self.stream):
            out, st_out = _enqueue_uniform(blob, n_rows, low, span, row_lo, row_hi, dev)
            if self.back is None:
                self.back = torch.empty(_RNG_WORDS + 2, dtype=torch.int32, pin_memory=True)
            back = self.back  # one speculation in flight at a time: take() or the next launch() has consumed the last
            back.copy_(st_out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.pending = dict(key=(int(n_rows), tuple(low.tolist()), tuple(span.tolist()), int(row_lo), int(row_hi), _dev_index(dev)),
                            blob=blob, out=out, st_out=st_out, back=back, ev=ev)

    def take(self, n_rows, low, span, row_lo, row_hi, dev):
        p, self.pending = self.pending, None
        if p is None:
            return None
        key = (int(n_rows), tuple(low.tolist()), tuple(span.tolist()), int(row_lo), int(row_hi), _dev_index(dev))
        if p["key"] != key or not torch.equal(torch.get_rng_state(), p["blob"]):
            self.misses += 1
            p["ev"].synchronize()  # let the speculative kernel finish before its buffers go back to the allocator
            return None
        import numpy as np
        p["ev"].synchronize()
        main = torch.cuda.current_stream(dev)
        main.wait_event(p["ev"])
        p["out"].record_stream(main)
        _rng_set_from(p["blob"].numpy(), p["back"].numpy().view(np.uint32))
        self.hits += 1
        return p["out"]


def device_uniform(n_rows, low, high, row_lo=0, row_hi=None, device=None, prefetch=None):
    """``Uniform(low, high).sample((n_rows,))`` of torch's CPU default generator, drawn ON THE DEVICE from the
    generator's own state (klerg_mt19937_uniform): bit-exact samples, and the host generator is left where the host
    draw would have left it.  Returns rows [row_lo, row_hi) as a CUDA tensor [rows, D].  ``prefetch``: a
    UniformPrefetch that may already hold exactly this draw."""
    import numpy as np
    dev = device or _dev()
    low = torch.as_tensor(low, dtype=torch.float32).reshape(-1)
    span = torch.as_tensor(high, dtype=torch.float32).reshape(-1) - low  # fp32, as Uniform.rsample computes it
    row_hi = n_rows if row_hi is None else row_hi
    if prefetch is not None:
        got = prefetch.take(n_rows, low, span, row_lo, row_hi, dev)
        if got is not None:
            return got
    blob = torch.get_rng_state()
    out, st_out = _enqueue_uniform(blob, n_rows, low, span, row_lo, row_hi, dev)
    back = st_out.cpu().numpy().view(np.uint32)  # one small D2H: the host generator continues from here
    _rng_set_from(blob.numpy(), back)
    return out


def gather_rows(table, idx):
    """table [cap,S] float32, idx [M] int64 (device) -> [M,S]."""
    M, S = idx.numel(), table.shape[1]
    out = torch.empty((M, S), dtype=torch.float32, device=table.device)
    if M:
        cabi.check(cabi.load().klerg_gather_rows(cabi.ptr(table), S, cabi.ptr(idx), M, cabi.ptr(out),
                                                 cabi.stream_ptr()), "klerg_gather_rows")
    return out

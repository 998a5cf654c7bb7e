#!/usr/bin/env python
"""Benchmark of the KL-ergodic hot path (BASELINE.json metric) on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host CPU

Metric: state-sample pairs/s (BASELINE.json "metric"); evals/s is printed beside it.
A *step* is one fused planner eval on one batch of synthetic input:
rollout + barrier -> footprint of the H planned states over the workspace samples
(+ cached history footprint) -> renormalise -> importance ratio -> gradient for all H
states -> adjoint -> du, djdlam, u*.  2*H*N pairs per eval, ONE kernel launch.

Workload: "c4" = configs[3] of BASELINE.json, the 1e7-sample 6-D end-effector-pose
workspace the north_star quotes its roofline / scaling target on (H=50, 1e5-state
memory buffer).  It fits one GPU; with N GPUs the SAME 1e7 samples are sharded
(strong scaling) and the totals / gradient partials cross NVLink inside the kernel.
The other configs are reachable with --workload (c2 numbers ride along under "also").

`value`     inputs resident in HBM; K evals replayed from a CUDA graph over >= 2
            independent input sets (each larger than L2); timed with CUDA events.
`e2e`       the same metric through the reference-shaped public API (`Robot.step()`):
            host RNG samples -> H2D, target density, full planner step, D2H of the plan.
`roofline`  the fused eval kernel (the only kernel in the timed region), against the
            FP32-pipe / MUFU peaks measured on this box with the library's probes.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "embodied-active-learning-vision_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402,F401
import torch  # noqa: E402

import workloads as wl  # noqa: E402

L2_BYTES = 126 * 2 ** 20
METRIC = "klerg_state_sample_pairs_per_s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c4", choices=list(wl.WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="override the workspace size (total samples)")
    ap.add_argument("--weak", action="store_true", help="--samples per GPU instead of sharding a fixed workspace")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary c2 measurement")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="serialise consecutive evals (no programmatic dependent launch overlap)")
    ap.add_argument("--profile-evals", type=int, default=0,
                    help="run this many eager evals inside cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of control_torch on the host cores
# --------------------------------------------------------------------------------------
def oracle_pairs(r, n, m, m_all):
    """pairs evaluated by one OracleRobot.step(): history + spread + evals."""
    H = r.horizon
    return m * n + m_all * n + r.n_cost_evals * H * n + r.n_grad_evals * 2 * H * n


def run_oracle_steps(name, n, m, steps, warmup, seed=0):
    from oracle import klerg_oracle as ko
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    torch.manual_seed(seed)
    w = wl.WORKLOADS[name]
    lims = [wl.LIMS[s] for s in w["states"]]
    target = wl.make_target(w["target"], lims, seed=1, device="cpu")
    r = ko.OracleRobot(**wl.robot_kwargs(name, target, n_samples=n, cap=max(m, 8)))
    r.test(min(n, 1000))
    for row in wl.random_walk_history(name, m):
        r.memory_buffer.push(row)
    pairs = evals = 0
    hist = spread = evp = 0
    t_total = 0.0
    for k in range(warmup + steps):
        r.n_cost_evals = r.n_grad_evals = 0
        m_all = len(r.memory_buffer)
        t0 = time.perf_counter()
        r.step(n, m, save_update=True)
        dt = time.perf_counter() - t0
        if k >= warmup:
            t_total += dt
            pairs += oracle_pairs(r, n, min(m, m_all), m_all)
            evals += r.n_cost_evals + r.n_grad_evals
            hist += min(m, m_all) * n
            spread += m_all * n
            evp += (r.n_cost_evals + 2 * r.n_grad_evals) * r.horizon * n
    return dict(pairs=pairs, evals=evals, seconds=t_total, cores=cores, steps=steps, mix=pair_mix(hist, spread, evp))


# bounded sample of every workload for the host CPU legs: the workload's own history size M (same pair mix as the GPU
# arm's Robot.step()), the sample count reduced so that a Robot.step() of the oracle port takes seconds, not hours
REF_SAMPLES = {"c1": 1_000, "c2": 20_000, "c3": 20_000, "c4": 4_000, "c5": 20_000}


def pair_mix(hist_pairs, spread_pairs, eval_pairs):
    tot = max(hist_pairs + spread_pairs + eval_pairs, 1)
    return {"history": hist_pairs / tot, "spread": spread_pairs / tot, "evals": eval_pairs / tot}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = wl.WORKLOADS[args.workload]
    n_full = args.samples or w["N"]
    m = w["M"]
    n = min(n_full, REF_SAMPLES[args.workload])
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # the whole run must end within a few minutes: probe one step, then time as many steps as fit ~150 s
    probe = run_oracle_steps(args.workload, n, m, 1, 0)
    t_step = probe["seconds"]
    budget = 150.0
    if t_step * (steps + warm) > budget:
        warm = min(warm, 1)
        steps = max(1, int(budget / t_step) - warm)
    res = run_oracle_steps(args.workload, n, m, steps, warm)
    value = res["pairs"] / res["seconds"]
    clamp = "" if steps == max(1, args.steps) else f" ({args.steps} steps requested; clamped to fit the time budget)"
    sample = (f"{steps}{clamp} full Robot.step() calls of the oracle port (torch CPU fp32, all host threads) at N={n} samples "
              f"(workload {args.workload} has N={n_full}) and the workload's own M={m}; the reference itself is Python and "
              f"cannot travel to the GPU box, the port is pinned to it by tests/golden")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": res["seconds"] / max(res["steps"], 1) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, n_total=n, n_local=n,
                                  note="host CPU, oracle port of control_torch; a step here is one whole Robot.step() "
                                       "(history + spread + ~13 evals) at a REDUCED sample count N (same M, same pair mix "
                                       "as the GPU arm's e2e Robot.step(); the GPU arm's `also.same_size` runs exactly this "
                                       "N and M)"),
        "evals_per_s": res["evals"] / res["seconds"], "pair_mix": res["mix"],
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(name, n_total, n_local, note="", m=None):
    w = wl.WORKLOADS[name]
    big = n_local * 4 * (len(w["states"]) + 3) > L2_BYTES
    return {"workload": f"{name}: states={w['states']} H={w['H']} N={n_total} (per GPU {n_local}) M={m if m is not None else w['M']} "
                        f"target={w['target']} barrier=on R=0.5 dt=0.2",
            "pairs_per_eval": 2 * w["H"] * n_total,
            "l2_policy": ("alternating independent input sets, each larger than the 126 MB L2" if big else
                          "ring of independent input sets, together larger than the 126 MB L2"),
            "note": note}


# --------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [s for s, p in zip(sm, power) if p > 250] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
    return world, rank, local, pg


def time_events(fn, reps):
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(reps):
        fn()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / reps  # ms


def measure_peaks(cabi, sms):
    """Lane-op rates measured on this box (roofline denominators): scalar FFMA, packed FFMA2, MUFU.EX2."""
    lib = cabi.load()
    out = torch.zeros(4, device="cuda")
    res = {}
    for kind, name, lanes in ((0, "ffma_per_s", 1), (2, "ffma2_lane_ops_per_s", 2), (1, "ex2_per_s", 1)):
        iters, blocks = 2000, sms * 8
        f = lambda: cabi.check(lib.klerg_peak_probe(kind, iters, blocks, cabi.ptr(out), cabi.stream_ptr()), "peak")  # noqa: E731
        f()
        ms = min(time_events(f, 3) for _ in range(3))
        res[name] = blocks * 256 * iters * 64 * lanes / (ms * 1e-3)
    res["fp32_lane_ops_per_s"] = max(res["ffma_per_s"], res["ffma2_lane_ops_per_s"])
    return res


def eval_roofline(D, H, n_local, peaks, hbm_gbs):
    """Seconds one fused eval needs at the roofline, with the minimal instruction mix of a pair
    (DESIGN.md): forward D FADD + 1 FMUL + (D-1) FFMA + 1 FADD = 2D+1 lane-ops + 1 MUFU.EX2;
    gradient D FADD + 1 FMUL + (D-1) FFMA + 1 FMUL + D FFMA = 3D+1 lane-ops + 1 MUFU.EX2."""
    pairs = H * n_local
    fp32, mufu = peaks["fp32_lane_ops_per_s"], peaks["ex2_per_s"]
    t_fwd = max(pairs * (2 * D + 1) / fp32, pairs / mufu)
    t_grad = max(pairs * (3 * D + 1) / fp32, pairs / mufu)
    bytes_fwd = n_local * 4 * (D + 2)        # samples + q_base in, q out
    bytes_grad = n_local * 4 * (D + 2)       # samples + q + p in
    t_hbm = (bytes_fwd + bytes_grad) / (hbm_gbs * 1e9)
    bound = "fp32+mufu"
    t = t_fwd + t_grad
    if t_hbm > t:
        t, bound = t_hbm, "hbm"
    flops = pairs * (4 * D + 1) + pairs * (6 * D + 1)  # algorithmic flop per pair, SURVEY.md 8(d)
    return dict(seconds=t, bound=bound, flops=flops, bytes=bytes_fwd + bytes_grad,
                fwd_bound="mufu" if pairs / mufu > pairs * (2 * D + 1) / fp32 else "fp32",
                grad_bound="mufu" if pairs / mufu > pairs * (3 * D + 1) / fp32 else "fp32")


def target_decoder_block(dev, timed):
    """p(s) of the VAE sensor model (vae/vae.py:244-275) on the tensor cores: klerg_target_decoder_pdf at the c2
    and c4 workspace sizes, against the tf32 peak (half the measured bf16 peak) and the torch-CPU arithmetic."""
    from control_torch.target_decoder import DeviceTarget
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16, src = 2250.0, "nominal 2.25 PF bf16 (B200_PROFILING.md fallback)"
    if os.path.exists(peaks_file):
        bf16, src = json.load(open(peaks_file))["bf16_tflops"], "MEASURED_PEAKS.json bf16_tflops (burst)"
    out = {}
    for tag, n_t, sdim in (("c2", 100_000, 3), ("c4_size", 10_000_000, 6)):
        model = wl.SyntheticVAETarget(sdim, seed=1)
        devt = DeviceTarget(model, dev)
        smp = (torch.rand(n_t, sdim, device=dev) * 2.3 - 1.15).contiguous()
        devt.refresh()
        t_kernel = timed(lambda: devt.evaluate_packed(smp), 5)
        t_call = timed(lambda: devt.pdf_torch(smp), 5)
        devt.check_fault()
        zd, h1, h2 = 16, 256, 512
        flop_alg = 2.0 * n_t * ((zd + sdim) * h1 + h1 * h2 + h2)
        flop_issued = 3 * 2.0 * n_t * h1 * h2
        out[tag] = {"samples": n_t, "s_dim": sdim, "ms_kernel": t_kernel * 1e3, "ms_pdf_torch_call": t_call * 1e3,
                    "samples_per_s": n_t / t_kernel,
                    # frac: against the NOMINAL dense tf32 peak (1.1 PFLOP/s); MEASURED_PEAKS.json has no tf32 entry, and the
                    # figure derived from its bf16 number (half of a power-capped burst) is not a measured tf32 peak.  The
                    # counter that says how busy the pipe is: ncu sm__pipe_tensor_cycles_active 82.5 % (profiles/r01_ncu_full_target_decoder_1e7.csv)
                    "roofline": {"bound": "tensor", "achieved": flop_issued / t_kernel / 1e12, "peak": 1100.0, "unit": "TFLOP/s",
                                 "frac": flop_issued / t_kernel / 1e12 / 1100.0,
                                 "ncu_pipe_tensor_cycles_active": 0.825,
                                 "algorithmic_tflops": flop_alg / t_kernel / 1e12,
                                 "derived_peak_from_bf16": bf16 / 2, "frac_of_derived": flop_issued / t_kernel / 1e12 / (bf16 / 2),
                                 "peak_basis": f"tf32 dense = half of {src}; achieved counts the 3 tf32 MMAs of the 3xTF32 "
                                               "split (fp32-accurate result); algorithmic = 2*N*(19*256+256*512+512) fp32 flop"}}
        del smp
    model = wl.SyntheticVAETarget(3, seed=1)
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    x = torch.rand(100_000, 3) * 2.3 - 1.15
    model.pdf_torch(x[:1000])
    t0 = time.perf_counter()
    model.pdf_torch(x)
    t_cpu = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": 100_000 / t_cpu, "unit": "samples/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                           "sample": "pdf_torch of the same decoder with the reference's torch operators on 1e5 samples"}
    out["note"] = ("ms_pdf_torch_call = the call a controller step makes (weights re-read from the model + H2D + pack + "
                   "kernel); the L2 is cold for the 1e7-sample input (160 MB in/out)")
    return out


def build_sets(name, n_total, rank, group, dev, engine, Robot, PlannerContext, max_sets=96):
    """Independent input sets (samples, p, q_base, u) resident in HBM; every rank holds its slice."""
    w = wl.WORKLOADS[name]
    st = w["states"]
    D, H = len(st), w["H"]
    lims = [wl.LIMS[s] for s in st]
    target = wl.make_target(w["target"], lims, seed=1, device=dev)
    kw = wl.robot_kwargs(name, target, n_samples=n_total)
    torch.manual_seed(1234)
    probe = Robot(process_group=None, **kw)  # only used to build specs (dyn, barrier, limits)
    lo_i, hi_i = group.shard_bounds(n_total)
    n = hi_i - lo_i
    bytes_per_set = engine.padded(n) * 4 * (D + 3)
    n_sets = min(max_sets, max(2, int(L2_BYTES * 1.5 / bytes_per_set) + 1))
    hist = wl.random_walk_history(name, w["M"], seed=0).to(dev)
    lo = torch.tensor([a for a, _ in lims]) * 1.15
    hi = torch.tensor([b for _, b in lims]) * 1.15
    x0 = torch.tensor(kw["x0"], dtype=torch.float32, device=dev)
    sets = []
    for s in range(n_sets):
        ctx = PlannerContext(probe.planner.spec, probe.barrier.spec(), probe.explr_locs.tolist(), H,
                             torch.diagonal(probe.R_inv).tolist(), probe.control_lim[:, 0].tolist(),
                             probe.control_lim[:, 1].tolist(), alpha=1.0, group=group)
        g = torch.Generator(device=dev).manual_seed(100 + 17 * s + rank)
        smp = lo.to(dev) + torch.rand(n, D, generator=g, device=dev) * (hi - lo).to(dev)
        ctx.set_samples(smp, probe.std.tolist(), 1.0, n_total=n_total)
        ctx.set_state(x0)
        p_raw = torch.cat([target.pdf_torch(c) for c in smp.split(1_000_000)]).contiguous()
        p, p_stats, _ = engine.target_weight(2, smp, lo.tolist(), hi.tolist(), None, p_raw, n_total, 1.0, True, group)
        ctx.set_target(p, p_stats)
        ctx.set_history(hist)
        ctx.samples = None  # raw samples are not read by an eval
        ctx.buf.v_costs = None  # candidate scratch is not needed for gradient evals
        ctx.u = wl.random_controls((H, D), seed=1000 + s).to(dev)
        sets.append(ctx)
        del smp, p_raw
    torch.cuda.synchronize()
    return dict(sets=sets, n=n, n_sets=n_sets, bytes_per_set=bytes_per_set, D=D, H=H, kw=kw, probe=probe, target=target)


def timed_evals(args, S, K, warmup, world, rank, lib, dev):
    """K fused evals over the ring of input sets: CUDA-graph replay, CUDA events, max over ranks."""
    sets, n_sets = S["sets"], S["n_sets"]

    def eval_on(i):
        c = sets[i % n_sets]
        return c.gradient(c.u)

    from control_torch import engine as _engine
    for i in range(max(warmup, 3)):
        eval_on(i)
    torch.cuda.synchronize()
    # the timed evals are independent of one another (separate input sets and outputs): let the next eval's CTAs
    # start while the previous eval's last CTA is still in its gather + adjoint tail
    _engine.set_eval_overlap(not args.no_overlap)
    if args.profile_evals:
        torch.cuda.profiler.start()
        for i in range(args.profile_evals):
            eval_on(i)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    use_graph = not args.no_graph
    graphs, reps, rem = [], 0, 0
    launches_per_eval = 1.0
    if use_graph:
        try:
            cyc = min(n_sets * 4, K)
            reps, rem = divmod(K, cyc)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):  # allocate this stream's workspace and warm it outside the capture
                    eval_on(i)
                side.synchronize()
                c0 = lib.klerg_launch_count()
                for length in ([cyc] + ([rem] if rem else [])):
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=side):
                        keep = [eval_on(i) for i in range(length)]
                    graphs.append((gr, length, keep))
                launches_per_eval = (lib.klerg_launch_count() - c0) / (cyc + rem)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for gr, _, _ in graphs:  # warm the instantiated graphs
                gr.replay()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({e!r}); timing eager launches", file=sys.stderr)
            use_graph, graphs = False, []
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = lib.klerg_launch_count()
    start.record()
    if use_graph:
        for _ in range(reps):
            graphs[0][0].replay()
        if rem:
            graphs[1][0].replay()
    else:
        for i in range(K):
            eval_on(i)
    end.record()
    torch.cuda.synchronize()
    ms_total = start.elapsed_time(end)
    gpu_launches = int(round(launches_per_eval * K)) if use_graph else int(lib.klerg_launch_count() - launches1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        dist.barrier()
    _engine.set_eval_overlap(False)
    return dict(ms_total=ms_total, gpu_launches=gpu_launches, use_graph=use_graph)


def cuda_arm(args):
    from control_torch import _cabi as cabi
    from control_torch import engine
    from control_torch.klerg import Robot
    from control_torch.planner import PlannerContext

    lib = cabi.load()
    world, rank, local, pg = dist_setup(args)
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    group = engine.ShardGroup(pg)
    dev = torch.device("cuda", local)
    sms = C.c_int()
    lib.klerg_device_info(C.byref(sms), None, None)
    sms = sms.value
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_gbs, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_file):
        hbm_gbs, hbm_src = json.load(open(peaks_file))["hbm_gbs"], "measured"

    w = wl.WORKLOADS[args.workload]
    n_total = (args.samples or w["N"]) * (world if args.weak else 1)
    sampler = ClockSampler(local) if rank == 0 else None
    S = build_sets(args.workload, n_total, rank, group, dev, engine, Robot, PlannerContext)
    D, H, n = S["D"], S["H"], S["n"]
    K = args.steps
    # sharded runs: every rank must hold bit-identical results (they drive identical host control flow)
    rank_identical = None
    if world > 1:
        import torch.distributed as dist
        c0 = S["sets"][0]
        g0 = c0.gradient(c0.u)
        flat = torch.cat([g0["du"].reshape(-1), g0["djdlam"].reshape(-1), g0["u_star"].reshape(-1)]).contiguous()
        every = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(every, flat)
        rank_identical = all(torch.equal(every[0], e) for e in every)
        if not rank_identical:
            raise SystemExit("ranks disagree on du / djdlam / u* of the same eval: the measurement is void")
    res = timed_evals(args, S, K, args.warmup, world, rank, lib, dev)
    if res is None:
        return
    ms_total = res["ms_total"]
    pairs_per_eval = 2 * H * n_total
    value = pairs_per_eval * K / (ms_total * 1e-3)
    evals_per_s = K / (ms_total * 1e-3)
    clocks = sampler.stop() if sampler else None
    if engine.fused_fault():
        raise SystemExit("fused eval reported a meeting-point fault; the measurement is void")

    # ---- roofline of the fused eval kernel (the only kernel launched in the timed region) ------------
    roof, peaks = None, None
    if rank == 0:
        peaks = measure_peaks(cabi, sms)
        rf = eval_roofline(D, H, n, peaks, hbm_gbs)
        t_launch = ms_total / K * 1e-3
        traffic = None
        try:  # measured once per round with ncu --set full (not re-measured live: ncu cannot run inside the bench)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(f"{args.workload}:{n}")
            traffic = tr["bytes"] if tr else None
        except (OSError, ValueError):
            pass
        roof = {"kernel": "eval_grad_kernel (rollout + footprint + renormalize + gradient + adjoint, one launch per eval)",
                "bound": "fp32" if rf["bound"] != "hbm" else "hbm",
                "achieved": rf["flops"] / t_launch / 1e12, "peak": rf["flops"] / rf["seconds"] / 1e12, "unit": "TFLOP/s",
                "frac": rf["seconds"] / t_launch, "traffic": traffic, "algorithmic_bytes": rf["bytes"],
                "launch_us": t_launch * 1e6, "roofline_us": rf["seconds"] * 1e6,
                "peak_basis": (f"FP32-pipe and MUFU peaks measured on this box with the library's probes: packed FFMA2 "
                               f"{peaks['ffma2_lane_ops_per_s']:.3e} lane-ops/s, scalar FFMA {peaks['ffma_per_s']:.3e}/s, "
                               f"MUFU.EX2 {peaks['ex2_per_s']:.3e}/s; time = H*N*max((2D+1)/fp32, 1/mufu) forward "
                               f"[{rf['fwd_bound']}-bound] + H*N*max((3D+1)/fp32, 1/mufu) gradient [{rf['grad_bound']}-bound]; "
                               f"achieved/peak in algorithmic flop ((4D+1)+(6D+1) per state-sample pair, SURVEY 8d). "
                               f"MEASURED_PEAKS.json holds only HBM and bf16 tensor peaks, neither of which bounds this kernel."),
                "hbm": {"achieved_gbs": rf["bytes"] / t_launch / 1e9, "peak_gbs": hbm_gbs, "peak_source": hbm_src,
                        "frac": rf["bytes"] / t_launch / 1e9 / hbm_gbs},
                "measured_peaks": peaks}

    # ---- secondary workload: c2 (configs[1]) evals, same protocol, short -----------------------------
    also = None
    if not args.no_also and args.workload != "c2":
        n2 = wl.WORKLOADS["c2"]["N"]
        S2 = build_sets("c2", n2, rank, group, dev, engine, Robot, PlannerContext)
        r2 = timed_evals(args, S2, 2000, 20, world, rank, lib, dev)
        if rank == 0:
            rf2 = eval_roofline(S2["D"], S2["H"], S2["n"], peaks, hbm_gbs)
            t2 = r2["ms_total"] / 2000 * 1e-3
            also = {"c2": {"workload": workload_config("c2", n2, S2["n"])["workload"],
                           "pairs_per_s": 2 * S2["H"] * n2 / t2, "evals_per_s": 1.0 / t2,
                           "us_per_eval": t2 * 1e6, "roofline_frac": rf2["seconds"] / t2,
                           "note": "1e5 samples leave ~650 samples per SM: the eval is latency-bound (launch, two grid "
                                   "meeting points, serial rollout/adjoint), not throughput-bound"}}
        del S2

    # ---- the other BASELINE configs, single GPU, short: c1 eval, c3 batched candidates, c5 belief targets ----
    if also is not None and world == 1 and rank == 0:
        def timed(fn, reps):
            for _ in range(2):
                fn()
            return time_events(fn, reps) * 1e-3

        def robot_step_ms(name, steps=40):
            """Robot.step() at the workload's own N and M through the public API (wall clock, one D2H per step)."""
            w_ = wl.WORKLOADS[name]
            tgt = wl.make_target(w_["target"], [wl.LIMS[s] for s in w_["states"]], seed=1, device=dev)
            torch.manual_seed(7)
            rb = Robot(process_group=None, **wl.robot_kwargs(name, tgt))
            rb.test(1000)
            for row in wl.random_walk_history(name, w_["M"], seed=5):
                rb.memory_buffer.push(row)
            for _ in range(5):
                rb.step(w_["N"], w_["M"], save_update=True)
            torch.cuda.synchronize()
            t0_ = time.perf_counter()
            for _ in range(steps):
                rb.step(w_["N"], w_["M"], save_update=True)
            torch.cuda.synchronize()
            out = {"ms_per_robot_step": (time.perf_counter() - t0_) / steps * 1e3, "robot_steps_timed": steps,
                   "planner_loop": "device (klerg_plan_optimize)" if rb.device_loop else "host"}
            if not args.no_cpu:
                ro = run_oracle_steps(name, w_["N"], w_["M"], 2, 1)
                out["cpu_oracle_ms_per_robot_step"] = ro["seconds"] / ro["steps"] * 1e3
                out["cpu_oracle_cores"] = ro["cores"]
            return out

        # the dominant kernel of the e2e step: history footprint + spread of the memory buffer, ONE M x N pass, squared
        # distances on the tensor cores (klerg_footprint_sum_max_tc) - its own bound is the MUFU pipe (one exp per pair)
        if S["sets"] and args.workload == "c4":
            c0 = S["sets"][0]
            hist4 = wl.random_walk_history(args.workload, w["M"], seed=0).to(dev)
            pairs4 = float(hist4.shape[0]) * c0.n
            t_tc = timed(lambda: engine.footprint_sum_max(c0.spec, hist4, hist4.shape[0], c0.packed, c0.n, tensor_cores=True), 2)
            t_cc = timed(lambda: engine.footprint_sum_max(c0.spec, hist4, hist4.shape[0], c0.packed, c0.n, tensor_cores=False), 2)
            also["history_pass"] = {
                "workload": f"{args.workload}: {hist4.shape[0]} buffer rows x {c0.n} samples, sum + max in one pass",
                "ms_per_pass": t_tc * 1e3, "pairs_per_s": pairs4 / t_tc, "ms_cuda_core_pass": t_cc * 1e3,
                "roofline": {"bound": "mufu", "achieved": pairs4 / t_tc, "peak": peaks["ex2_per_s"], "unit": "exp/s",
                             "frac": pairs4 / t_tc / peaks["ex2_per_s"]},
                "note": "e_ij as one K = 8 tcgen05.mma kind::tf32 step (3xTF32), min / exp / add on the CUDA cores; the "
                        "CUDA-core pass (D + 4 lane-ops per pair, FP32-pipe bound) beside it"}
            del hist4, c0

        S1 = build_sets("c1", wl.WORKLOADS["c1"]["N"], rank, group, dev, engine, Robot, PlannerContext, max_sets=8)
        r1 = timed_evals(args, S1, 2000, 20, world, rank, lib, dev)
        also["c1"] = {"workload": workload_config("c1", wl.WORKLOADS["c1"]["N"], S1["n"])["workload"],
                      "us_per_eval": r1["ms_total"] / 2000 * 1e3, "evals_per_s": 2000 / (r1["ms_total"] * 1e-3)}
        del S1
        also["c1"].update(robot_step_ms("c1"))
        if "c2" in also:
            also["c2"].update(robot_step_ms("c2"))
        S3 = build_sets("c3", wl.WORKLOADS["c3"]["N"], rank, group, dev, engine, Robot, PlannerContext, max_sets=2)
        c3 = S3["sets"][0]
        c3.buf.v_costs = torch.empty((c3.buf.max_g, c3.buf.ld), dtype=torch.float32, device=dev)
        gU = torch.Generator().manual_seed(2)
        U3 = (c3.u.cpu().unsqueeze(0) + 0.1 * torch.randn(1024, S3["H"], S3["D"], generator=gU)).to(dev)
        l0 = lib.klerg_launch_count()
        c3.costs(U3)
        launches_c3 = int(lib.klerg_launch_count() - l0)
        t3 = timed(lambda: c3.costs(U3), 3)
        also["c3"] = {"workload": workload_config("c3", wl.WORKLOADS["c3"]["N"], S3["n"])["workload"] + " B=1024 candidates",
                      "ms_per_batch": t3 * 1e3, "pairs_per_s": 1024 * S3["H"] * S3["n"] / t3, "candidates_per_s": 1024 / t3,
                      "launches_per_batch": launches_c3,
                      "roofline_frac": 1024 * S3["H"] * S3["n"] * max((2 * S3["D"] + 1) / peaks["fp32_lane_ops_per_s"],
                                                                    1 / peaks["ex2_per_s"]) / t3,
                      "note": "get_cost of 1024 candidates: 8 candidates per fused launch, forward pair pass only"}
        lims5 = [wl.LIMS[c] for c in wl.WORKLOADS["c5"]["states"]]
        n5 = S3["n"]
        smp5 = torch.rand(n5, S3["D"], device=dev) * 2.3 - 1.15
        P5 = torch.stack([wl.make_target("gmm", lims5, seed=20 + k, device=dev).pdf_torch(smp5) for k in range(16)])
        st5 = torch.stack([engine.vector_stats(P5[k].contiguous())[:1] for k in range(16)])
        c3.set_targets(P5.contiguous(), st5)
        paths = {}
        for path in ("fused", "tensor"):
            c3.targets_path = path
            paths[path] = timed(lambda: c3.gradient_targets(c3.u, check=False), 5)
        if engine.targets_gradient_fault():
            raise SystemExit("klerg_kl_gradient_targets reported a timed-out wait; the measurement is void")
        del c3.targets_path  # back to the class default
        tensor_default = c3.targets_path != "fused" and c3._tensor_targets_ok(16)
        t5 = paths["tensor"] if tensor_default else paths["fused"]
        also["c5"] = {"workload": "c5: states=xyz H=50 N=1000000 K=16 belief targets", "ms_per_16_target_gradient": t5 * 1e3,
                      "target_gradients_per_s": 16 / t5, "state_sample_target_triples_per_s": 16 * S3["H"] * n5 / t5,
                      "ms_fused_launch_pair_pass_per_target": paths["fused"] * 1e3,
                      "ms_shared_psi_tensor_core": paths["tensor"] * 1e3, "default_path": "tensor" if tensor_default else "fused",
                      "note": "fused: one launch - shared rollout / forward pass / q, per-target gradient pair pass + adjoint; "
                              "tensor: rollout, forward pass, klerg_kl_gradient_targets (psi once per pair, tcgen05 tf32 3xTF32 "
                              "contraction over the samples for all 16 targets), one adjoint launch"}
        del S3, c3, U3, P5, smp5
        torch.cuda.empty_cache()
        also["target_decoder"] = target_decoder_block(dev, timed)

    # ---- with several GPUs: the same eval with the per-GPU workspace held fixed (weak scaling) -------
    if world > 1 and not args.weak and not args.no_also:
        del S["sets"][:]
        torch.cuda.empty_cache()
        Sw = build_sets(args.workload, (args.samples or w["N"]) * world, rank, group, dev, engine, Robot, PlannerContext)
        rw = timed_evals(args, Sw, 100, 5, world, rank, lib, dev)
        if rank == 0:
            tw = rw["ms_total"] / 100 * 1e-3
            also = dict(also or {})
            also["weak_scaling"] = {"workload": workload_config(args.workload, (args.samples or w["N"]) * world, Sw["n"])["workload"],
                                    "pairs_per_s": 2 * H * (args.samples or w["N"]) * world / tw, "us_per_eval": tw * 1e6,
                                    "note": "per-GPU samples held at the single-GPU workload; the headline value above "
                                            "shards the fixed 1e7-sample workspace (strong scaling)"}
        del Sw
        torch.cuda.empty_cache()

    # ---- e2e: Robot.step() through the public API with host buffers ---------------------------------
    e2e = None
    same_size = None
    if not args.no_e2e:
        kw, m = S["kw"], w["M"]
        del S["sets"][:]  # free the resident sets; the planner owns its buffers
        torch.cuda.empty_cache()

        def time_robot_steps(n_steps, n_warm, n_samples, device_rng, kwargs):
            """Robot.step() through the public API, wall clock around the timed steps (max over ranks)."""
            torch.manual_seed(7)
            robot = Robot(process_group=pg, **kwargs)
            robot.device_rng = device_rng
            robot.test(1000)
            for row in wl.random_walk_history(args.workload, min(m, robot.memory_buffer.capacity), seed=5):
                robot.memory_buffer.push(row)
            hist = spread = 0
            for k in range(n_warm + n_steps):
                if k == n_warm:
                    torch.cuda.synchronize()
                    if world > 1:
                        import torch.distributed as dist
                        dist.barrier()
                    t0 = time.perf_counter()
                    c0_, g0_ = robot.stats["cost_evals"], robot.stats["grad_evals"]
                m_all = len(robot.memory_buffer)
                robot.step(n_samples, m, save_update=True)
                if k >= n_warm:
                    hist += min(m, m_all) * n_samples
                    spread += m_all * n_samples
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            if world > 1:
                import torch.distributed as dist
                tt = torch.tensor([t], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = float(tt.item())
            ce, ge = robot.stats["cost_evals"] - c0_, robot.stats["grad_evals"] - g0_
            evp = (ce + 2 * ge) * H * n_samples
            wrapped = getattr(robot, "_wrapped_target", None)
            on_dev = device_rng and robot._device_draw_ok()
            del robot
            torch.cuda.empty_cache()
            return dict(seconds=t, ce=ce, ge=ge, pairs=hist + spread + evp, mix=pair_mix(hist, spread, evp), wrapped=wrapped,
                        device_draw=on_dev, m_all=m_all)

        steps_e = args.e2e_steps
        r = time_robot_steps(steps_e, 2, n_total, True, kw)
        ce, ge, m_all = r["ce"], r["ge"], r["m_all"]
        # bytes that cross PCIe per step, from the tensors the step copies
        n_loc = n
        per_eval_h2d, per_eval_d2h = H * D * 4, 0
        sample_bytes = (624 * 4 + 16) if r["device_draw"] else n_loc * D * 4  # generator state vs the samples themselves
        p_bytes = 0 if r["device_draw"] else (n_loc * 4 if r["wrapped"] is None else r["wrapped"]._staging.numel() * 4)
        h2d = sample_bytes + p_bytes + min(m, m_all) * 8 + (m_all * 8 if m < m_all else 0) + (ce + ge) // steps_e * per_eval_h2d + 2 * D * 4
        d2h = (626 * 4 if r["device_draw"] else 0) + (ge // steps_e) * (H * 4 + H * D * 4 + 4) + (ce // steps_e) * 4 * 9 \
            + (H + 1) * 2 * D * 4 + 2 * D * 4
        e2e = {"value": r["pairs"] / r["seconds"], "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "control_torch.klerg.Robot.step(num_target_samples=N, num_traj_samples=M, save_update=True)",
               "steps": steps_e, "ms_per_robot_step": r["seconds"] / steps_e * 1e3, "evals_per_s": (ce + ge) / r["seconds"],
               "evals_per_robot_step": (ce + ge) / steps_e, "pair_mix": r["mix"],
               "note": ("the step's host inputs are the robot state, the controls and torch's CPU generator: the workspace "
                        "samples are drawn ON THE DEVICE from the generator's state (bit-exact with the host draw, "
                        "klerg_mt19937_uniform; the draw of step k+1 runs one step ahead on a side stream and is used only "
                        "if the generator is found where it was left), the host draws the memory-buffer permutation and "
                        "copies the indices; "
                        if r["device_draw"] else
                        "samples drawn by the host torch RNG (reference order) and copied H2D every step; ")
                       + "the target density is evaluated on the device; history footprint and spread are ONE M_all x N pass "
                         "(squared distances on the tensor cores: klerg_footprint_sum_max_tc); "
                         "pair_mix: most pairs of a step are history / spread pairs (forward-only, cheaper than eval pairs), "
                         "which is why e2e pairs/s can exceed the eval-only `value`"}
        if world == 1 and rank == 0:
            rh = time_robot_steps(min(3, steps_e), 1, n_total, False, kw)
            e2e["host_draw"] = {"ms_per_robot_step": rh["seconds"] / min(3, steps_e) * 1e3,
                                "h2d_bytes_per_step": int(n_loc * D * 4 + min(m, m_all) * 8),
                                "note": "the same steps with the samples drawn by the host generator and copied H2D (4*N*D bytes)"}
            # the reference arm's exact problem: reduced N, the workload's own M (same_config comparison)
            n_ref = min(n_total, REF_SAMPLES[args.workload])
            kw_ref = wl.robot_kwargs(args.workload, S["target"], n_samples=n_ref)
            rs = time_robot_steps(10, 2, n_ref, True, kw_ref)
            same_size = {"workload": workload_config(args.workload, n_ref, n_ref)["workload"],
                         "ms_per_robot_step": rs["seconds"] / 10 * 1e3, "pairs_per_s": rs["pairs"] / rs["seconds"],
                         "pair_mix": rs["mix"],
                         "note": "Robot.step() at the N and M `bench.py --impl reference` runs: same problem on both arms"}

    # ---- cpu_baseline: the oracle port on this box's host cores (rank 0, N=1 only) ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = min(n_total, REF_SAMPLES[args.workload])  # ~10-30 s of host work incl. the warm-up step
        m_cpu = w["M"]
        r_cpu = run_oracle_steps(args.workload, n_cpu, m_cpu, 2, 1)
        cpu = {"value": r_cpu["pairs"] / r_cpu["seconds"], "unit": "pairs/s", "cores": r_cpu["cores"], "kind": "port",
               "sample": f"2 OracleRobot.step() calls (torch CPU fp32, all host threads) at N={n_cpu}, M={m_cpu}, "
                         f"H={H}: {r_cpu['evals']} evals in {r_cpu['seconds']:.2f} s",
               "evals_per_s": r_cpu["evals"] / r_cpu["seconds"], "ms_per_robot_step": r_cpu["seconds"] / 2 * 1e3,
               "pair_mix": r_cpu["mix"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, n_total, n,
                                      note=("CUDA graph replay" if res["use_graph"] else "eager launches")
                                      + f", {S['n_sets']} input sets x {S['bytes_per_set'] / 2**20:.1f} MiB per GPU"),
            "evals_per_s": evals_per_s, "gpu_launches": res["gpu_launches"], "clocks": clocks, "e2e": e2e,
            "roofline": roof, "cpu_baseline": cpu, "also": also,
        }
        if same_size is not None:
            line["also"] = dict(line["also"] or {})
            line["also"]["same_size"] = same_size
        if rank_identical is not None:
            line["rank_identical"] = rank_identical
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()

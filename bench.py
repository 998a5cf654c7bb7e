#!/usr/bin/env python
"""Benchmark of the KL-ergodic hot path (BASELINE.json metric) on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host CPU

Metric: state-sample pairs/s (BASELINE.json "metric"); evals/s is printed beside it.
A *step* is one planner eval on one batch of synthetic input:
rollout + barrier -> footprint of the H planned states over the N workspace samples
(+ cached history footprint) -> renormalise -> KL -> importance ratio -> gradient for
all H states -> adjoint -> du, djdlam, u*.  2*H*N pairs per eval.

Workload (N=1): config[1] of BASELINE.json = "c2": 3-D xyz workspace, horizon 50, 1e5
samples, 3000 history rows, p from a random-init VAE-style uncertainty head.  With
N GPUs every rank owns 1e5 samples of an N*1e5-sample workspace (weak scaling) and
the per-eval totals / gradient partials cross NVLink in two small all-gathers.

`value`     inputs resident in HBM; K evals replayed from a CUDA graph over a ring of
            independent input sets larger than L2; timed with CUDA events.
`e2e`       the same metric through the reference-shaped public API (`Robot.step()`):
            host RNG samples -> H2D, target density, full planner step, D2H of the plan.
`roofline`  dominant kernel alone, CUDA events, against the measured FP32/MUFU issue peaks.
"""
import argparse
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

import workloads as wl  # noqa: E402

L2_BYTES = 126 * 2 ** 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(wl.WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="override samples per GPU")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
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
    return dict(pairs=pairs, evals=evals, seconds=t_total, cores=cores, steps=steps)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = wl.WORKLOADS[args.workload]
    n_full = args.samples or w["N"]
    m = min(w["M"], 3000)
    # bounded sample: probe one step at reduced size, then pick N so K steps fit ~150 s
    probe_n = min(n_full, 10_000)
    probe = run_oracle_steps(args.workload, probe_n, m, 1, 1)
    rate = probe["pairs"] / probe["seconds"]
    per_step_pairs_full = probe["pairs"] * (n_full / probe_n)
    budget = 150.0
    n = n_full
    if per_step_pairs_full / rate * (args.steps + args.warmup) > budget:
        n = int(max(1000, min(n_full, n_full * budget / (per_step_pairs_full / rate * (args.steps + args.warmup)))))
    res = run_oracle_steps(args.workload, n, m, args.steps, args.warmup)
    value = res["pairs"] / res["seconds"]
    sample = (f"{args.steps} full Robot.step() calls of the oracle port (torch CPU fp32) at N={n} samples "
              f"(workload {args.workload} has N={n_full}), M={m}")
    line = {
        "impl": "reference", "metric": "klerg_state_sample_pairs_per_s", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["seconds"] / max(res["steps"], 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_per_gpu=n, note="host CPU, oracle port of control_torch"),
        "evals_per_s": res["evals"] / res["seconds"],
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n_per_gpu, note=""):
    w = wl.WORKLOADS[args.workload]
    return {"workload": f"{args.workload}: states={w['states']} H={w['H']} N={n_per_gpu}/GPU M={w['M']} "
                        f"target={w['target']} barrier=on R=0.5 dt=0.2",
            "pairs_per_eval": 2 * w["H"] * n_per_gpu, "l2_policy": "ring of independent input sets > 126 MB L2",
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


def graph_time(fn, reps):
    """ms per call of `fn` with launch gaps removed: `reps` calls captured in one CUDA graph."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        with torch.cuda.graph(gr, stream=side):
            keep = [fn() for _ in range(reps)]
    torch.cuda.current_stream().wait_stream(side)
    gr.replay()
    ms = min(time_events(gr.replay, 2) for _ in range(3)) / reps
    del keep
    return ms


def measure_peaks(cabi, sms):
    """FFMA and MUFU.EX2 issue rates (ops/s) measured on this box: roofline denominators."""
    lib = cabi.load()
    out = torch.zeros(4, device="cuda")
    res = {}
    for kind, name in ((0, "ffma_per_s"), (1, "ex2_per_s")):
        iters, blocks = 2000, sms * 8
        f = lambda: cabi.check(lib.klerg_peak_probe(kind, iters, blocks, cabi.ptr(out), cabi.stream_ptr()), "peak")
        f()
        ms = min(time_events(f, 3) for _ in range(3))
        res[name] = blocks * 256 * iters * 64 / (ms * 1e-3)
    return res


def roofline_time(D, kind, pairs, peaks, bytes_moved, hbm_gbs):
    """Seconds at the issue roofline: minimal instruction mix of one pair
    forward: D FADD + D FFMA + 1 FADD on the FP32 pipes + 1 MUFU.EX2; gradient: + 1 FMUL + D FFMA."""
    fp32 = (2 * D + 1) if kind == "forward" else (3 * D + 1)
    t_fp32 = pairs * fp32 / peaks["ffma_per_s"]
    t_mufu = pairs * 1 / peaks["ex2_per_s"]
    t_issue = pairs * (fp32 + 1) / peaks["ffma_per_s"]  # one issue slot per warp instruction, 4 schedulers/SM
    t_hbm = bytes_moved / (hbm_gbs * 1e9)
    terms = {"fp32": t_fp32, "mufu": t_mufu, "issue": t_issue, "hbm": t_hbm}
    bound = max(terms, key=terms.get)
    return terms[bound], bound, terms


def cuda_arm(args):
    from control_torch import _cabi as cabi
    from control_torch import engine
    from control_torch.klerg import Robot
    from control_torch.planner import PlannerContext

    cabi.load()
    world, rank, local, pg = dist_setup(args)
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    group = engine.ShardGroup(pg)
    dev = torch.device("cuda", local)
    lib = cabi.load()
    import ctypes as C
    sms = C.c_int()
    lib.klerg_device_info(C.byref(sms), None, None)
    sms = sms.value

    w = wl.WORKLOADS[args.workload]
    st = w["states"]
    D, H = len(st), w["H"]
    n = args.samples or w["N"]  # per GPU (weak scaling)
    m = w["M"]
    lims = [wl.LIMS[s] for s in st]
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_gbs, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_file):
        hbm_gbs, hbm_src = json.load(open(peaks_file))["hbm_gbs"], "measured"

    # ---- inputs: a ring of independent (samples, p, q_base, u) sets larger than L2 ------------
    target = wl.make_target(w["target"], lims, seed=1, device=dev)
    kw = wl.robot_kwargs(args.workload, target, n_samples=n * world)
    torch.manual_seed(1234 + rank)
    probe = Robot(process_group=None, **kw)  # only used to build specs (dyn, barrier, limits)
    bytes_per_set = engine.padded(n) * 4 * (D + 3)
    n_sets = min(96, max(2, int(L2_BYTES * 1.5 / bytes_per_set) + 1))
    hist = wl.random_walk_history(args.workload, m, seed=rank).to(dev)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    lo = torch.tensor([a for a, _ in lims]) * 1.15
    hi = torch.tensor([b for _, b in lims]) * 1.15
    sets = []
    x0 = torch.tensor(kw["x0"], dtype=torch.float32, device=dev)
    for s in range(n_sets):
        ctx = PlannerContext(probe.planner.spec, probe.barrier.spec(), probe.explr_locs.tolist(), H,
                             torch.diagonal(probe.R_inv).tolist(), probe.control_lim[:, 0].tolist(),
                             probe.control_lim[:, 1].tolist(), alpha=1.0, group=group)
        smp = (lo + torch.rand(n, D, generator=g) * (hi - lo)).to(dev)
        ctx.set_samples(smp, probe.std.tolist(), 1.0)
        ctx.set_state(x0)
        p_raw = target.pdf_torch(smp).contiguous()
        p, p_stats, _ = engine.target_weight(2, smp, lo.tolist(), hi.tolist(), None, p_raw, n * world, 1.0, True, group)
        ctx.set_target(p, p_stats)
        ctx.set_history(hist)
        ctx.samples = None  # raw samples are not read by an eval
        ctx.u = wl.random_controls((H, D), seed=1000 * rank + s).to(dev)
        sets.append(ctx)
    del smp, p_raw
    torch.cuda.synchronize()

    def eval_on(i):
        c = sets[i % n_sets]
        return c.gradient(c.u)

    # ---- warm-up (eager) ------------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        eval_on(i)
    torch.cuda.synchronize()

    if args.profile_evals:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for i in range(args.profile_evals):
            eval_on(i)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # ---- timed region: exactly K evals ----------------------------------------------------------
    launches0 = lib.klerg_launch_count()
    use_graph = not args.no_graph
    graphs = []
    K = args.steps
    if use_graph:
        try:
            cyc = min(n_sets, K)
            reps, rem = divmod(K, cyc)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                engine.workspace(1)
                for length in ([cyc] + ([rem] if rem else [])):
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=side):
                        keep = [eval_on(i) for i in range(length)]
                    graphs.append((gr, length, keep))
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            launches_per_eval = (lib.klerg_launch_count() - launches0) / (cyc + rem)
            for gr, _, _ in graphs:  # warm the instantiated graphs
                gr.replay()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({e!r}); timing eager launches", file=sys.stderr)
            use_graph = False
            graphs = []
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
    if use_graph:
        gpu_launches = int(round(launches_per_eval * K))
    else:
        gpu_launches = int(lib.klerg_launch_count() - launches1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        dist.barrier()
    pairs_per_eval = 2 * H * n * world
    value = pairs_per_eval * K / (ms_total * 1e-3)
    evals_per_s = K / (ms_total * 1e-3)

    # ---- per-kernel timing and roofline (rank 0's view; single-GPU kernels) -----------------------
    kernels = {}
    roof = None
    if rank == 0:
        peaks = measure_peaks(cabi, sms)
        c0 = sets[0]
        ro = engine.rollout(c0.dyn, c0.bar, c0.x0, c0.u, want_lin=True)
        pre = ro["traj"][0][:H].contiguous()
        outs = {}

        def k_fwd(i=[0]):
            c = sets[i[0] % n_sets]
            i[0] += 1
            outs["v"], outs["tot"] = engine.footprint(c.spec, 0, pre, c.packed, c.n, add_in=c.q_base)
            return outs["v"]

        def k_grad(i=[0]):
            c = sets[i[0] % n_sets]
            i[0] += 1
            return engine.kl_gradient_fused(c.spec, pre, c.packed, c.n, outs["v"][0], outs["tot"].unsqueeze(0), c.p)

        k_fwd()
        k_grad()
        torch.cuda.synchronize()
        reps_k = n_sets
        for name, fn, kind in (("footprint_kernel", k_fwd, "forward"), ("grad_kernel", k_grad, "gradient")):
            ms = graph_time(fn, reps_k)
            pairs = H * n
            bytes_moved = n * 4 * (D + 2) if kind == "forward" else n * 4 * (D + 2)
            t_roof, bound, terms = roofline_time(D, kind, pairs, peaks, bytes_moved, hbm_gbs)
            flops_pair = (4 * D + 1) if kind == "forward" else (6 * D + 1)
            kernels[name] = {"ms": ms, "pairs_per_s": pairs / (ms * 1e-3), "frac_of_issue_roofline": t_roof / (ms * 1e-3),
                             "bound": bound, "algorithmic_tflops": pairs * flops_pair / (ms * 1e-3) / 1e12,
                             "hbm_gbs": bytes_moved / (ms * 1e-3) / 1e9}
        dom = max(kernels, key=lambda k: kernels[k]["ms"])
        kd = kernels[dom]
        kind = "forward" if dom == "footprint_kernel" else "gradient"
        flops_pair = (4 * D + 1) if kind == "forward" else (6 * D + 1)
        peak_tflops = kd["algorithmic_tflops"] / kd["frac_of_issue_roofline"]
        roof = {"kernel": dom, "bound": "fp32+mufu issue (" + kd["bound"] + ")", "achieved": kd["algorithmic_tflops"],
                "peak": peak_tflops, "unit": "TFLOP/s", "frac": kd["frac_of_issue_roofline"], "traffic": None,
                "peak_basis": (f"pairs/s at the tighter of FP32-pipe, MUFU and issue-slot limits, measured on this box: "
                               f"FFMA {peaks['ffma_per_s']:.3e}/s, MUFU.EX2 {peaks['ex2_per_s']:.3e}/s; x {flops_pair} "
                               f"algorithmic flop/pair (SURVEY 8d)"),
                "hbm": {"achieved_gbs": kd["hbm_gbs"], "peak_gbs": hbm_gbs, "peak_source": hbm_src,
                        "frac": kd["hbm_gbs"] / hbm_gbs},
                "measured_peaks": peaks}

    # ---- e2e: Robot.step() through the public API with host buffers ---------------------------------
    e2e = None
    if not args.no_e2e:
        torch.manual_seed(7)
        robot = Robot(process_group=pg, **kw)
        robot.test(1000)
        for row in wl.random_walk_history(args.workload, min(m, robot.memory_buffer.capacity), seed=5):
            robot.memory_buffer.push(row)
        n_tot = n * world
        pairs = 0
        for k in range(3 + args.e2e_steps):
            if k == 3:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                c0_, g0_ = robot.stats["cost_evals"], robot.stats["grad_evals"]
                pairs = 0
            m_all = len(robot.memory_buffer)
            robot.step(n_tot, m, save_update=True)
            if k >= 3:
                pairs += min(m, m_all) * n_tot + m_all * n_tot
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        ce, ge = robot.stats["cost_evals"] - c0_, robot.stats["grad_evals"] - g0_
        pairs += (ce + 2 * ge) * H * n_tot
        steps_e = args.e2e_steps
        h2d = n_tot // world * D * 4 + min(m, m_all) * 8 + (ce + ge) // steps_e * H * D * 4 + 2 * D * 4
        d2h = (ge // steps_e) * (H * 4 + H * D * 4) + (ce // steps_e) * 4 + (H + 1) * 2 * D * 4 + 2 * D * 4
        e2e = {"value": pairs / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "control_torch.klerg.Robot.step(num_target_samples=N, num_traj_samples=M, save_update=True)",
               "steps": steps_e, "ms_per_robot_step": t_e2e / steps_e * 1e3, "evals_per_s": (ce + ge) / t_e2e,
               "evals_per_robot_step": (ce + ge) / steps_e, "target_device": str(dev)}

    clocks = sampler.stop() if sampler else None

    # ---- cpu_baseline: the oracle port on this box's host cores (rank 0, N=1 only) ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = min(n, 20_000)
        res = run_oracle_steps(args.workload, n_cpu, min(m, 3000), 2, 1)
        cpu = {"value": res["pairs"] / res["seconds"], "unit": "pairs/s", "cores": res["cores"], "kind": "port",
               "sample": f"2 OracleRobot.step() calls (torch CPU fp32, all host threads) at N={n_cpu}, M={min(m, 3000)}, "
                         f"H={H}: {res['evals']} evals in {res['seconds']:.2f} s",
               "evals_per_s": res["evals"] / res["seconds"]}

    if rank == 0:
        line = {
            "metric": "klerg_state_sample_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n, note=("CUDA graph replay" if use_graph else "eager launches")
                                      + f", {n_sets} input sets x {bytes_per_set / 2**20:.1f} MiB"),
            "evals_per_s": evals_per_s, "gpu_launches": gpu_launches, "clocks": clocks, "e2e": e2e,
            "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()

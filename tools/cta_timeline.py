"""Per-CTA wall-clock timeline of one fused gradient eval (instrumented build: KLERG_VARIANT=_stamps).

  KLERG_VARIANT=_stamps python tools/cta_timeline.py c4 1250000
  KLERG_VARIANT=_stamps python -m torch.distributed.run --nproc-per-node 8 ... tools/cta_timeline.py c4     # sharded

Under torchrun the workspace of N samples is sharded over the ranks (the evals exchange over NVLink) and rank 0 and the
last rank print their timelines.

Prints, over the CTAs of the launch, when each phase ended relative to the first CTA's start (min / median / max, us):
how long the slowest CTA keeps everybody waiting at the two meeting points, and what the finisher CTA adds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
import numpy as np
import torch
import workloads as wl
from control_torch import engine
from control_torch.klerg import Robot
from control_torch.planner import PlannerContext

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else wl.WORKLOADS[name]["N"]
w = wl.WORKLOADS[name]
lims = [wl.LIMS[s] for s in w["states"]]
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
group = engine.SINGLE
if world > 1:
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = engine.ShardGroup(dist.group.WORLD)
dev = torch.device("cuda", torch.cuda.current_device())
n_total = n
s_lo, s_hi = group.shard_bounds(n_total)
target = wl.make_target("gmm", lims, seed=1, device=dev)
kw = wl.robot_kwargs(name, target, n_samples=n)
probe = Robot(**kw)
D, H = len(w["states"]), w["H"]
ctx = PlannerContext(probe.planner.spec, probe.barrier.spec(), probe.explr_locs.tolist(), H,
                     torch.diagonal(probe.R_inv).tolist(), probe.control_lim[:, 0].tolist(),
                     probe.control_lim[:, 1].tolist(), alpha=1.0, group=group)
g = torch.Generator(device=dev).manual_seed(0)
lo = (torch.tensor([a for a, _ in lims]) * 1.15).to(dev)
hi = (torch.tensor([b for _, b in lims]) * 1.15).to(dev)
smp = (lo + torch.rand(n_total, D, generator=g, device=dev) * (hi - lo))[s_lo:s_hi].contiguous()
n = s_hi - s_lo
ctx.set_samples(smp, probe.std.tolist(), 1.0, n_total=n_total)
ctx.set_state(torch.tensor(kw["x0"], dtype=torch.float32, device=dev))
p_raw = torch.cat([target.pdf_torch(c) for c in smp.split(1_000_000)]).contiguous()
p, p_stats, _ = engine.target_weight(2, smp, lo.tolist(), hi.tolist(), None, p_raw, n_total, 1.0, True, group)
ctx.set_target(p, p_stats)
ctx.set_history(wl.random_walk_history(name, min(w["M"], 3000)).to(dev))
u = wl.random_controls((H, D), seed=3).to(dev)
for _ in range(6):
    ctx.gradient(u)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    if rank not in (0, world - 1):
        dist.destroy_process_group()
        sys.exit(0)
    if rank != 0:
        import time
        time.sleep(1.0)  # keep the two printouts apart
st = engine.debug_cta_stamps(ctx.peers if world > 1 else None).numpy().astype(np.int64)
nb = int((st[:, 0] > 0).sum())
st = st[:nb]
t0 = st[:, 0].min()
names = ["start", "rollout", "forward", "meet1", "gradient", "entry sums"]
print(f"{name} N={n_total} over {world} rank(s), rank {rank} holds {n}: {nb} CTAs; times in us since the first CTA started (min / median / max over CTAs)")
for i, nm in enumerate(names):
    c = (st[:, i] - t0) / 1e3
    print(f"  {nm:12s} {c.min():9.2f} {np.median(c):9.2f} {c.max():9.2f}")
d = np.diff(st[:, :6], axis=1) / 1e3
for i, nm in enumerate(["rollout", "forward", "meet1 (wait+exchange)", "gradient", "entry sums (wait+sum)"]):
    print(f"  duration {nm:24s} min {d[:, i].min():8.2f} median {np.median(d[:, i]):8.2f} max {d[:, i].max():8.2f}")
fin = st[nb - 1]
print(f"  finisher CTA: end at {(fin[6] - t0) / 1e3:.2f} us, its tail (entry sums -> end) {(fin[6] - fin[5]) / 1e3:.2f} us")
cyc = engine.debug_stamps()
names2 = ["rollout(states)", "forward", "meet1", "gradient", "row sums", "lin+gather", "adjoint+end"]
print("  finisher phase cycles:", {k: cyc[i + 1] - cyc[i] for i, k in enumerate(names2)}, "| stage u/x0", cyc[8], "rollout_block", cyc[9] - cyc[8])

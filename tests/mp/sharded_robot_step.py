"""Sample-sharded Robot.step() with a VAE-like target vs the single-GPU controller (run under torchrun, one rank per
GPU).  Every rank draws the same host samples, keeps its slice, evaluates the target density of ITS slice with the
tensor-core decoder (p is point-wise, no exchange) and takes part in the in-kernel NVLink exchanges of the fused evals.
All ranks must return the same action, and rank 0 re-runs the whole sequence on one GPU for comparison."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200"), os.path.join(ROOT, "tests", "golden")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cases import ROBOT_CASES, TARGET_CASES, DecoderModel, robot_kwargs, seed_buffer_states  # noqa: E402
from control_torch.klerg import Robot  # noqa: E402


def run(process_group, steps=3):
    rc = ROBOT_CASES["xyz_small"]
    case = dict(TARGET_CASES["default"])
    torch.manual_seed(7)
    model = DecoderModel(case, torch.randn(1, case["zd"], generator=torch.Generator().manual_seed(2)))
    r = Robot(process_group=process_group, **robot_kwargs(rc, model))
    r.test(rc["n"])
    for s in seed_buffer_states(r.robot.state, rc):
        r.memory_buffer.push(s)
    n = 4099  # ragged against the 2-rank split and the 128-row decoder tiles
    out = [r.step(n, rc["m"], save_update=True) for _ in range(steps)]
    assert r._wrapped_target.stats["evals"] >= steps
    return out, r.u.clone()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res, u = run(dist.group.WORLD)
    # every rank holds the same controls (the exchanges combine in rank order -> bit-identical host control flow)
    flat = torch.cat([torch.as_tensor(np.concatenate([np.ravel(x) for x in step])) for step in res] + [u.reshape(-1)]).cuda().double()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        ref, u_ref = run(None)
        for (sa, va, ca), (sb, vb, cb) in zip(ref, res):
            np.testing.assert_allclose(sb, sa, rtol=1e-3, atol=1e-5)
            np.testing.assert_allclose(cb, ca, rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(u.numpy(), u_ref.numpy(), rtol=1e-3, atol=1e-5)
        print("SHARDED_ROBOT OK" if ok else "SHARDED_ROBOT ranks disagree", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

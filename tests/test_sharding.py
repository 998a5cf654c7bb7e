"""Sample sharding: host-side partition / gather logic on CPU with gloo
(world_size 2), and the NVLink-mailbox fused evals on >= 2 GPUs when the box has them."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
    from control_torch import engine
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = engine.ShardGroup(dist.group.WORLD)
        assert (g.world, g.rank) == (world, rank)
        # contiguous, disjoint, covering slices - also for ragged totals
        for n in (0, 1, 7, 100_003):
            lo, hi = g.shard_bounds(n)
            spans = [None] * world
            dist.all_gather_object(spans, (lo, hi))
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
        # rank-ordered gather of per-rank partial blocks (what the unfused path reduces on the device)
        block = torch.tensor([[float(rank), 10.0 + rank]], dtype=torch.float64)
        out = g.gather_blocks(block)
        assert out.shape == (world, 1, 2)
        assert out[:, 0, 0].tolist() == [float(r) for r in range(world)]
        assert g.peers is not None
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_group_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_rank_group_has_no_peers():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
    from control_torch import engine
    assert engine.SINGLE.world == 1 and engine.SINGLE.peers() is None
    assert engine.SINGLE.shard_bounds(10) == (0, 10)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("c2", 100_003), ("c1", 1_000)])
def test_sharded_fused_evals_match_single_gpu(name, n):
    """2 ranks, one per GPU: totals and gradient partials cross NVLink inside the fused kernels."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29800 + os.getpid() % 100
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp", "sharded_parity.py"), name, str(n)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SHARDED_PARITY OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
def test_sharded_robot_step_with_model_target():
    """2 ranks: Robot.step() with a VAE-like target (tensor-core decoder on each rank's sample slice) against the
    single-GPU controller."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29900 + os.getpid() % 100
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp", "sharded_robot_step.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SHARDED_ROBOT OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]

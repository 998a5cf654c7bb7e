"""klerg_mt19937_uniform: the workspace-sample draw of Robot.get_samples continued on the device from torch's CPU
generator.  Bit-exact against the host draw (and the numpy oracle), generator state handed back correctly, sharded row
ranges, and whole Robot.step() sequences identical with the host draw."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import workloads as wl  # noqa: E402
from oracle import mt19937_oracle as mo  # noqa: E402


@pytest.mark.parametrize("seed,warm,n,D", [(0, 0, 1000, 3), (5, 17, 7, 2), (11, 3, 624 * 3 + 5, 6), (3, 623, 250_003, 6),
                                           (9, 624, 1, 1), (2, 1, 0, 4), (4, 100, 2_000_000, 3)])
def test_device_draw_bit_exact_with_host(seed, warm, n, D):
    from control_torch import engine
    low = torch.tensor([-1.15, -1.15, 2.1, -0.9, -2.3, 0.3][:D])
    high = low + torch.tensor([2.3, 2.3, 1.7, 1.8, 4.6, 0.2][:D])
    torch.manual_seed(seed)
    if warm:
        torch.rand(warm)
    blob = torch.get_rng_state().numpy().tobytes()
    want = torch.distributions.Uniform(low, high).sample((n,))
    want_perm = torch.randperm(1234)
    torch.set_rng_state(torch.tensor(list(blob), dtype=torch.uint8))
    got = engine.device_uniform(n, low, high)
    assert got.shape == (n, D)
    assert torch.equal(got.cpu(), want.reshape(n, D))
    assert torch.equal(torch.randperm(1234), want_perm)  # the host generator continues where the host draw would have
    ref, _ = mo.uniform_samples(blob, n, low.numpy(), high.numpy())
    assert np.array_equal(got.cpu().numpy(), ref)
    # a rank's row range: the whole stream is generated, only its rows are written
    if n > 10:
        torch.set_rng_state(torch.tensor(list(blob), dtype=torch.uint8))
        part = engine.device_uniform(n, low, high, n // 3, n - 2)
        assert torch.equal(part.cpu(), want[n // 3: n - 2])
        assert torch.equal(torch.randperm(1234), want_perm)


def test_robot_sequences_identical_with_device_draw():
    """Robot.step() with the samples drawn on the device = the same controller with the host draw, bit for bit."""
    from control_torch.klerg import Robot
    name = "c2"
    w = wl.WORKLOADS[name]
    lims = [wl.LIMS[s] for s in w["states"]]
    out = []
    for device_rng in (False, True):
        target = wl.make_target("gmm", lims, seed=1, device="cuda")
        torch.manual_seed(3)
        r = Robot(**wl.robot_kwargs(name, target, n_samples=20_011, horizon=20, cap=64))
        r.device_rng = device_rng
        r.test(500)
        for row in wl.random_walk_history(name, 40, seed=5):
            r.memory_buffer.push(row)
        res = [r.step(20_011, 24, save_update=True) for _ in range(4)]
        out.append((res, r.u.clone(), r.last_hist_idx.clone(), torch.get_rng_state().clone()))
    (ra, ua, ia, sa), (rb, ub, ib, sb) = out
    for (x0, v0, c0), (x1, v1, c1) in zip(ra, rb):
        assert np.array_equal(x0, x1) and np.array_equal(v0, v1) and np.array_equal(c0, c1)
    assert torch.equal(ua, ub) and torch.equal(ia, ib) and torch.equal(sa, sb)
    assert r._prefetch is not None and r._prefetch.hits == 3 and r._prefetch.misses == 0  # steps 2..4 found their draw waiting


def test_speculative_draw_hits_and_misses():
    """engine.UniformPrefetch: the next draw enqueued ahead of time is used only if the generator is found exactly where
    the speculation started and the same draw is asked for; either way samples and generator equal the host draw."""
    from control_torch import engine
    low, high = torch.tensor([-1.0, 0.0, 2.0]), torch.tensor([1.0, 0.5, 4.0])
    dev = torch.device("cuda")
    pf = engine.UniformPrefetch()

    def host(n):
        state = torch.get_rng_state()
        want = torch.distributions.Uniform(low, high).sample((n,))
        after = torch.get_rng_state()
        torch.set_rng_state(state)
        return want, after

    torch.manual_seed(21)
    # hit
    want, after = host(5000)
    pf.launch(5000, low, high, 0, 5000, dev)
    got = engine.device_uniform(5000, low, high, prefetch=pf)
    assert (pf.hits, pf.misses) == (1, 0)
    assert torch.equal(got.cpu(), want) and torch.equal(torch.get_rng_state(), after)
    # somebody else consumed the generator in between: the speculation is dropped, the in-line draw runs
    pf.launch(5000, low, high, 0, 5000, dev)
    torch.rand(3)
    want, after = host(5000)
    got = engine.device_uniform(5000, low, high, prefetch=pf)
    assert (pf.hits, pf.misses) == (1, 1)
    assert torch.equal(got.cpu(), want) and torch.equal(torch.get_rng_state(), after)
    # a different row count, a different row range
    for args in ((4999, 0, 4999), (5000, 100, 4000)):
        pf.launch(5000, low, high, 0, 5000, dev)
        want, after = host(args[0])
        got = engine.device_uniform(args[0], low, high, args[1], args[2], prefetch=pf)
        assert torch.equal(got.cpu(), want[args[1]:args[2]]) and torch.equal(torch.get_rng_state(), after)
    assert (pf.hits, pf.misses) == (1, 3)
    # nothing pending
    want, after = host(10)
    assert torch.equal(engine.device_uniform(10, low, high, prefetch=pf).cpu(), want)
    assert (pf.hits, pf.misses) == (1, 3)

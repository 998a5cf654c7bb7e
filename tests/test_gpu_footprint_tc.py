"""klerg_footprint_sum_max_tc - the history footprint + spread pass with the squared distances on the tensor cores
(one K = 8 tf32 MMA step per pair block, 3xTF32) - against the CUDA-core pass and the oracle: ragged sizes around the
128-sample tiles and the 128-row chunks, empty / full summed part, D = 2, 3, 6, and the on-device fallback when the
states leave the radius of the expanded pair form."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
from oracle import klerg_oracle as ko  # noqa: E402


def _case(D, S, N, T, seed, std=0.2, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    samples = torch.rand(N, D, generator=g) * 2.3 - 1.15
    walk = torch.cumsum(0.03 * torch.randn(T, S, generator=g), 0)
    states = (walk - walk.mean(0)) * spread
    states = states.clamp(-1.0, 1.0)
    scale = torch.full((D,), std) * torch.tensor([1.0, -1.0] * 4)[:D]  # the sign of std must not matter
    return samples, states.contiguous(), scale


def _flag(engine):
    from control_torch import _cabi as cabi
    sc = engine._tc_scratch[(torch.cuda.current_device(), cabi.raw_stream())]
    return int(sc[32:36].view(torch.int32).item())


@pytest.mark.parametrize("D,S,N,T,T_sum", [(6, 12, 5000, 700, 300), (6, 12, 128, 256, 256), (3, 6, 1000, 513, 0), (6, 12, 150_000, 300, 200),
                                           (2, 4, 777, 1000, 1000), (6, 12, 4097, 2000, 1500), (4, 8, 130, 129, 1),
                                           (6, 12, 33_000, 3_000, 2_999)])
def test_tensor_core_pass_matches_cuda_cores_and_oracle(D, S, N, T, T_sum):
    from control_torch import _cabi as cabi, engine
    samples, states, scale = _case(D, S, N, T, seed=N + T)
    explr = list(range(D))
    spec = cabi.kernel_spec(D, S, explr, scale.tolist(), 1.0)
    dev = torch.device("cuda")
    packed = engine.pack_samples(spec, samples.to(dev))
    st = states.to(dev)
    s_tc, m_tc, tot_tc = engine.footprint_sum_max(spec, st, T_sum, packed, N, tensor_cores=True)
    torch.cuda.synchronize()
    assert _flag(engine) == 1, "the states lie within the radius: the tensor-core pass must have run"
    s_cc, m_cc, tot_cc = engine.footprint_sum_max(spec, st, T_sum, packed, N, tensor_cores=False)
    np.testing.assert_allclose(s_tc[:N].cpu().numpy(), s_cc[:N].cpu().numpy(), rtol=1e-4, atol=1e-37)
    np.testing.assert_allclose(m_tc[:N].cpu().numpy(), m_cc[:N].cpu().numpy(), rtol=1e-4, atol=1e-37)
    np.testing.assert_allclose(tot_tc.cpu().numpy(), tot_cc.cpu().numpy(), rtol=1e-5)
    want_s = ko.footprint_sum(states[:T_sum], samples, torch.tensor(explr), scale, torch.ones(1)) if T_sum else torch.zeros(N)
    want_m = ko.spread_max(states, samples, torch.tensor(explr), scale, 1.0)
    np.testing.assert_allclose(s_tc[:N].cpu().numpy(), want_s.numpy(), rtol=1e-4, atol=1e-30)
    np.testing.assert_allclose(m_tc[:N].cpu().numpy(), want_m.numpy(), rtol=1e-4, atol=1e-30)


def test_tensor_core_pass_falls_back_outside_the_radius():
    """A narrow kernel (std 0.01) puts the states far outside the expanded form's radius around any centre: the flag
    stays 0 and the CUDA-core pass behind the tensor-core launch produces the result - bit for bit its own."""
    from control_torch import _cabi as cabi, engine
    D, S, N, T, T_sum = 3, 6, 3000, 600, 400
    samples, states, scale = _case(D, S, N, T, seed=5, std=0.01, spread=3.0)
    spec = cabi.kernel_spec(D, S, list(range(D)), scale.tolist(), 1.0)
    dev = torch.device("cuda")
    packed = engine.pack_samples(spec, samples.to(dev))
    s_tc, m_tc, tot_tc = engine.footprint_sum_max(spec, states.to(dev), T_sum, packed, N, tensor_cores=True)
    torch.cuda.synchronize()
    assert _flag(engine) == 0
    s_cc, m_cc, tot_cc = engine.footprint_sum_max(spec, states.to(dev), T_sum, packed, N, tensor_cores=False)
    assert torch.equal(s_tc[:N], s_cc[:N]) and torch.equal(m_tc[:N], m_cc[:N]) and torch.equal(tot_tc, tot_cc)


def test_tensor_core_pass_validation():
    from control_torch import _cabi as cabi
    lib = cabi.load()
    assert lib.klerg_footprint_tc_scratch_bytes(1000) == 256 + (8 + 2) * 8192
    spec = cabi.kernel_spec(3, 6, [0, 1, 2], [0.1, 0.1, 0.1], 1.0)
    import ctypes as C
    assert lib.klerg_footprint_sum_max_tc(C.byref(spec), None, 10, 11, None, 100, 100, None, None, None, None, None, 0, None) != 0
    assert b"bad sizes" in lib.klerg_last_error()
    assert lib.klerg_footprint_sum_max_tc(C.byref(spec), None, 10, 5, None, 100, 100, None, None, None, None, None, 0, None) != 0
    assert b"null" in lib.klerg_last_error()


@pytest.mark.parametrize("D", [2, 3, 6])
def test_tensor_core_pass_at_the_edge_of_the_radius_vs_float64(D):
    """States up to |xc|^2 = 95 (in kernel widths) from the centre - the edge of what the radius guard admits - and
    samples a few widths from them: sum and max against a float64 evaluation of the same fp32 inputs, 1e-4."""
    import math
    from control_torch import _cabi as cabi, engine
    S, N, T = 2 * D, 2500, 1200
    g = torch.Generator().manual_seed(40 + D)
    std = 0.05
    a = math.sqrt(0.5 * math.log2(math.e) / std)  # scaled coordinate = a * x
    dirs = torch.randn(T, D, generator=g)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    radii = torch.sqrt(torch.rand(T, 1, generator=g)) * math.sqrt(95.0) / a
    radii[:2 * D] = math.sqrt(95.0) / a
    pos = dirs * radii
    for d in range(D):  # the bounding box is symmetric, so the centre is the origin and |xc|^2 reaches 95
        pos[2 * d] = 0.0
        pos[2 * d, d] = math.sqrt(95.0) / a
        pos[2 * d + 1] = -pos[2 * d]
    states = torch.zeros(T, S)
    states[:, :D] = pos
    samples = (pos[torch.randint(0, T, (N,), generator=g)] + torch.randn(N, D, generator=g) * (1.5 / a)).contiguous()
    spec = cabi.kernel_spec(D, S, list(range(D)), [std] * D, 1.0)
    dev = torch.device("cuda")
    packed = engine.pack_samples(spec, samples.to(dev))
    s_tc, m_tc, _ = engine.footprint_sum_max(spec, states.to(dev), T, packed, N, tensor_cores=True)
    torch.cuda.synchronize()
    assert _flag(engine) == 1
    d2 = ((samples.double().unsqueeze(1) - pos.double().unsqueeze(0)) ** 2).sum(2) / (2.0 * std)  # klerg_utils.py:7-10
    psi = torch.exp(-d2)
    np.testing.assert_allclose(s_tc[:N].cpu().double().numpy(), psi.sum(1).numpy(), rtol=1e-4)
    np.testing.assert_allclose(m_tc[:N].cpu().double().numpy(), psi.max(1).values.numpy(), rtol=1e-4)

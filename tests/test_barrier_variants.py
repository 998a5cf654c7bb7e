"""TiltBarrierFunction and VelocityBarrier (reference barrier.py:95-144, 162-205): oracle restatement against vectors
recorded from the live reference classes (CPU), and the B200 mirror classes against the same vectors (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
from oracle import klerg_oracle as ko  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "barrier_variants.npz")
ROBOT_RPW = torch.tensor([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]])
TRAY_RPW = torch.tensor([[2.39, 3.89], [-0.75, 0.75], [-2.0, 2.0]])


def _tilt_inputs(gold, tag):
    lim = torch.from_numpy(gold["tilt/lim"])
    xs = torch.from_numpy(gold["tilt/x"] if tag == "plain" else gold["tilt/x_mapped"])
    return lim, xs, (None if tag == "plain" else (ROBOT_RPW, TRAY_RPW))


@pytest.mark.parametrize("tag", ["plain", "mapped"])
def test_oracle_tilt_barrier_vs_reference_vectors(tag):
    gold = np.load(GOLD)
    lim, xs, amap = _tilt_inputs(gold, tag)
    lo, hi = lim[:, 0] + 0.1, lim[:, 1] - 0.1
    w, pw = torch.full((12,), 5.0), torch.full((12,), 4.0)
    w_lim = torch.from_numpy(gold[f"tilt/{tag}/w_b_lim"])
    vals, grads, tilt = [], [], None
    for x in xs:
        v, g, tilt = ko.tilt_barrier(x, lo, hi, w, pw, 3, 4, 5, w_lim, 2.45, ang_map=amap)
        vals.append(v)
        grads.append(g)
    np.testing.assert_allclose(torch.stack(vals).numpy(), gold[f"tilt/{tag}/value"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(torch.stack(grads).numpy(), gold[f"tilt/{tag}/grad"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose((tilt / torch.pi * w_lim).numpy(), gold[f"tilt/{tag}/w_lim_after"], rtol=1e-6)
    assert (gold[f"tilt/{tag}/value"] > 0).sum() > 10 and np.abs(gold[f"tilt/{tag}/grad"][:, 3:5]).max() > 0


def test_oracle_velocity_barrier_vs_reference_vectors():
    gold = np.load(GOLD)
    x_old, x_new = torch.from_numpy(gold["vel/x_old"]), torch.from_numpy(gold["vel/x_new"])
    band = torch.tile(torch.tensor([[-1.0, 1.0]]), (6, 1)) * 0.1
    skip = [s.lower() == s for s in "xyzXYZ"]
    out = [ko.velocity_barrier(a, b, band, 100.0, 4.0, skip) for a, b in zip(x_new, x_old)]
    np.testing.assert_allclose(torch.stack([o[0] for o in out]).numpy(), gold["vel/value"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(torch.stack([o[1] for o in out]).numpy(), gold["vel/grad"], rtol=1e-5, atol=1e-9)
    swapped = [ko.velocity_barrier(a, b, band, 100.0, 4.0, skip)[0] for a, b in zip(x_old, x_new)]  # __call__'s order
    np.testing.assert_allclose(torch.stack(swapped).numpy(), gold["vel/call"], rtol=1e-5, atol=1e-9)
    assert (gold["vel/value"] > 0).sum() > 5


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["plain", "mapped"])
def test_mirror_tilt_barrier_vs_reference_vectors(tag):
    from control_torch import barrier as kb
    from control_torch.klerg_utils import Lambda
    from franka.franka_utils import ws_conversion
    gold = np.load(GOLD)
    lim, xs, amap = _tilt_inputs(gold, tag)
    other = kb.BarrierFunction(b_lim=lim, barr_weight=5.0, b_buff=0.1, power=[4.0] * 12)
    fn = None if amap is None else Lambda(ws_conversion, amap)
    bar = kb.TiltBarrierFunction(other, "xyzrpw", tilt_lim=2.45, rot_to_angles_fn=fn)
    np.testing.assert_allclose(bar(xs).numpy(), gold[f"tilt/{tag}/value"], rtol=1e-4, atol=1e-5)
    got = torch.stack([bar.dbarr(x) for x in xs])
    np.testing.assert_allclose(got.numpy(), gold[f"tilt/{tag}/grad"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(other.b_lim[5].numpy(), gold[f"tilt/{tag}/w_lim_after"], rtol=1e-5)
    np.testing.assert_allclose(float(bar.barr(xs[7])), gold[f"tilt/{tag}/value"][7], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_mirror_velocity_barrier_vs_reference_vectors():
    from control_torch import barrier as kb
    gold = np.load(GOLD)
    x_old, x_new = torch.from_numpy(gold["vel/x_old"]), torch.from_numpy(gold["vel/x_new"])
    vb = kb.VelocityBarrier("xyzXYZ", b_lim=0.1, power=4, barr_weight=100.0)
    np.testing.assert_allclose(torch.stack([vb.barr(a, b) for a, b in zip(x_new[:12], x_old[:12])]).numpy(),
                               gold["vel/value"][:12], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(torch.stack([vb.dbarr(a, b) for a, b in zip(x_new[:12], x_old[:12])]).numpy(),
                               gold["vel/grad"][:12], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(vb(x_old, x_new).numpy(), gold["vel/call"], rtol=1e-4, atol=1e-8)

"""FingerprintDist.update_prior on the device (klerg_belief_update through the host mirror) against the vectors recorded
from the live reference (tests/golden/fingerprint_*.npz) and against the oracle on fresh inputs.  float64 like the
reference; tolerance 1e-11 relative (device exp/log differ from glibc's in the last bit)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_fingerprint import CASES, SUB, measurements  # noqa: E402
from oracle import fingerprint_oracle as fo  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-11


def _dist(case, capacity=64):
    from dist_modules.fingerprint_module import FingerprintDist
    return FingerprintDist(explr_states=case["states"], plot_idx=[0, 1], capacity=capacity,
                           lims=[list(x) for x in case["lims"]], thresh=case["thresh"], clip=case["clip"], name=("a", "b", "c"))


@pytest.mark.parametrize("name", list(CASES))
def test_belief_update_vs_reference_vectors(name):
    case = CASES[name]
    gold = np.load(os.path.join(HERE, "golden", f"fingerprint_{name}.npz"))
    fd = _dist(case)
    assert fd.grid.shape[0] == int(gold["grid_points"])
    np.testing.assert_array_equal(fd.grid[::SUB], gold["grid_rows"])
    np.testing.assert_array_equal(fd.lims, gold["lims_scaled"])
    assert fd.scale == float(gold["scale"])
    for k in range(int(gold["n_updates"])):
        locs, vals = gold[f"u{k}/locs"], gold[f"u{k}/vals"]
        if locs.shape[0] == 1:
            fd.push(locs[0], vals[0])
        else:
            fd.push_batch(locs, vals)
        np.testing.assert_allclose(fd.get_meas(separate=True)[1], gold[f"u{k}/processed"], rtol=1e-15)
        fd.update_prior()
        assert fd.position == 0 and not fd.full_buffer
        prior, prior_var = fd.prior, fd.prior_var
        np.testing.assert_allclose(prior[::SUB], gold[f"u{k}/prior"], rtol=RTOL, atol=0)
        np.testing.assert_allclose(prior_var[::SUB], gold[f"u{k}/prior_var"], rtol=RTOL, atol=0)
        np.testing.assert_allclose(prior.sum(), float(gold[f"u{k}/prior_sum"]), rtol=RTOL)
        np.testing.assert_allclose(prior_var.sum(), float(gold[f"u{k}/prior_var_sum"]), rtol=RTOL)
    assert fd.count == sum(int(gold[f"u{k}/locs"].shape[0]) for k in range(int(gold["n_updates"])))


def test_belief_update_vs_oracle_many_measurements():
    """xyzw grid (6.25e6 points would be the robot's; here 50^3) with 200 measurements in one batch, then the grid pdf."""
    case = dict(states="xyw", lims=[[-0.7, 0.9], [-1.0, 1.0], [-1.5, 2.0]], thresh=0.4, clip=1.5)
    fd = _dist(case, capacity=256)
    grid, lims, scale = fo.build_grid(case["lims"], case["states"])
    prior, prior_var = np.full(grid.shape[0], 0.5), np.full(grid.shape[0], 2.0)
    rng = np.random.default_rng(11)
    for n in (200, 3):
        locs, vals = measurements(case, n, rng)
        fd.push_batch(locs, vals)
        fd.update_prior()
        prior, prior_var = fo.update_prior(grid, prior, prior_var, locs, fo.process_meas(vals, case["thresh"], case["clip"]), scale)
        np.testing.assert_allclose(fd.prior, prior, rtol=RTOL, atol=0)
        np.testing.assert_allclose(fd.prior_var, prior_var, rtol=RTOL, atol=0)
    # pdf: uniform 0.5 until init, then the belief on the grid (inverted when asked), :591-606
    np.testing.assert_array_equal(fd.pdf(None, use_grid=True), np.full(grid.shape[0], 0.5))
    fd.init = True
    np.testing.assert_allclose(fd.pdf(None, use_grid=True), prior, rtol=RTOL)
    fd.invert = True
    np.testing.assert_allclose(fd.pdf(None, use_grid=True), -prior + prior.max() + prior.min(), rtol=1e-9)
    t = fd.pdf(None, use_grid=True, as_tensor=True)
    assert t.is_cuda and t.dtype.is_floating_point and t.shape[0] == grid.shape[0]


def test_belief_update_errors():
    from control_torch import _cabi as cabi
    case = CASES["xy"]
    fd = _dist(case)
    with pytest.raises(ValueError):
        fd.update_prior()          # no measurements, like the reference's get_meas
    fd.push(np.zeros(2), 0.3)
    with pytest.raises(NotImplementedError):
        fd.update_prior(smooth=True)
    lib = cabi.load()
    with pytest.raises(RuntimeError, match="null"):
        cabi.check(lib.klerg_belief_update(None, 10, 2, None, 1, 1.0, 0.0, None, None, None, None, None, None), "klerg_belief_update")

"""oracle/fingerprint_oracle.py against the vectors recorded from the live reference's FingerprintDist
(tests/golden/make_golden_fingerprint.py), and - when /root/reference is present - against the live class itself."""
import os

import numpy as np
import pytest

from oracle import fingerprint_oracle as fo

HERE = os.path.dirname(os.path.abspath(__file__))
import sys  # noqa: E402
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_fingerprint import CASES, SUB  # noqa: E402


@pytest.mark.parametrize("name", list(CASES))
def test_belief_update_vs_golden(name):
    case = CASES[name]
    gold = np.load(os.path.join(HERE, "golden", f"fingerprint_{name}.npz"))
    grid, lims, scale = fo.build_grid(case["lims"], case["states"])
    assert grid.shape[0] == int(gold["grid_points"])
    np.testing.assert_array_equal(grid[::SUB], gold["grid_rows"])
    np.testing.assert_array_equal(lims, gold["lims_scaled"])
    assert scale == float(gold["scale"])
    prior, prior_var = np.ones(grid.shape[0]) * 0.5, np.ones(grid.shape[0]) * 2.0  # init_uniform_grid, :541-544
    for k in range(int(gold["n_updates"])):
        locs, vals = gold[f"u{k}/locs"], gold[f"u{k}/vals"]
        proc = fo.process_meas(vals, case["thresh"], case["clip"])
        np.testing.assert_allclose(proc, gold[f"u{k}/processed"], rtol=1e-15, atol=0)
        np.testing.assert_allclose(fo.meas_footprint(locs, grid, scale / 2.0)[::SUB], gold[f"u{k}/meas_map"], rtol=1e-14, atol=1e-300)
        prior, prior_var = fo.update_prior(grid, prior, prior_var, locs, proc, scale)
        np.testing.assert_allclose(prior[::SUB], gold[f"u{k}/prior"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(prior_var[::SUB], gold[f"u{k}/prior_var"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(prior.sum(), float(gold[f"u{k}/prior_sum"]), rtol=1e-12)
        np.testing.assert_allclose(prior_var.sum(), float(gold[f"u{k}/prior_var_sum"]), rtol=1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference/franka_test/scripts"), reason="needs the reference checkout")
def test_belief_update_vs_live_reference():
    from make_golden_fingerprint import import_reference, measurements
    fm = import_reference()
    case = dict(states="xyw", lims=[[-0.7, 0.9], [-1.0, 1.0], [-1.5, 2.0]], thresh=0.4, clip=1.5)
    fd = fm.FingerprintDist(explr_states=case["states"], plot_idx=[0, 1], capacity=32, lims=[list(x) for x in case["lims"]],
                            thresh=case["thresh"], clip=case["clip"], name=("a", "b", "c"))
    grid, lims, scale = fo.build_grid(case["lims"], case["states"])
    np.testing.assert_array_equal(grid, fd.grid)
    prior, prior_var = np.ones(grid.shape[0]) * 0.5, np.ones(grid.shape[0]) * 2.0
    rng = np.random.default_rng(3)
    for n in (4, 1, 11):
        locs, vals = measurements(case, n, rng)
        fd.push_batch(locs, vals)
        proc = fd.get_meas(separate=True)[1].copy()
        fd.update_prior()
        prior, prior_var = fo.update_prior(grid, prior, prior_var, locs, fo.process_meas(vals, case["thresh"], case["clip"]), scale)
        np.testing.assert_allclose(fo.process_meas(vals, case["thresh"], case["clip"]), proc, rtol=1e-15)
        np.testing.assert_allclose(prior, fd.prior, rtol=1e-12, atol=0)
        np.testing.assert_allclose(prior_var, fd.prior_var, rtol=1e-12, atol=0)

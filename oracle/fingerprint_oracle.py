"""CPU restatement of the fingerprint belief update - TEST INFRASTRUCTURE ONLY (tests/ may import it; the product
path never does).  numpy float64 like the reference; every function cites the reference lines it follows
(/root/reference/franka_test/scripts/dist_modules/fingerprint_module.py, control/klerg_utils.py).  Pinned against
vectors recorded from the live reference: tests/golden/make_golden_fingerprint.py -> tests/golden/fingerprint_*.npz,
checked in tests/test_fingerprint_oracle.py."""
import numpy as np


def build_grid(lims, states, num_samples=50):
    """FingerprintDist.build_grid (fingerprint_module.py:504-519): yaw limits x1.33, all limits x1.15, 50 points per
    dimension, meshgrid order.  Returns (grid [50^D, D], scaled lims, default scale = 2.5 x largest spacing)."""
    lims = np.array(lims, dtype=np.float64).copy()
    if "w" in states:
        lims[states.rfind("w")] *= 1.33
    lims *= 1.15
    spacing = np.linspace(*lims.T, num_samples)
    mesh = np.meshgrid(*spacing.T)
    grid = np.c_[[m.ravel() for m in mesh]].T
    return grid, lims, float(np.max(spacing[1] - spacing[0]) * 2.5)


def process_meas(x, thresh, clip):
    """fingerprint_module.py:470-478: tanh of the thresholded measurement value."""
    if thresh is None:
        return x
    tmp = thresh - np.asarray(x, dtype=np.float64)
    tmp = tmp.copy()
    tmp[tmp > 0] /= thresh
    tmp[tmp < 0] /= (clip - thresh)
    return np.tanh(tmp)


def meas_footprint(locs, samples, std):
    """meas_footprint_vec (fingerprint_module.py:417-424): exp(-0.5 sum_d (loc_jd - s_gd)^2 / |std|), std >= 1e-6."""
    std = np.clip(std, 1e-6, None)
    inner = np.square(locs[None, :, :] - samples[:, None, :]) / np.abs(std)
    return np.exp(-0.5 * np.sum(inner, -1))


def renormalize(dist, dim=None, min_val=1e-6):
    """control/klerg_utils.py:41-54 (the numpy twin the fingerprint module imports)."""
    dist = dist / np.sum(dist, dim, keepdims=dim is not None)
    dist = np.log(np.clip(dist, min_val, None))
    dist = dist - np.max(dist, dim, keepdims=dim is not None)
    return np.exp(dist)


def update_prior(grid, prior, prior_var, locs, vals, scale):
    """FingerprintDist.update_prior (fingerprint_module.py:539-589, smooth=False, use_mask=False):
    (posterior, posterior_var) of the per-grid-point normal belief after n measurements (locs [n, D], processed
    values vals [n])."""
    n = locs.shape[0]
    meas_map = renormalize(meas_footprint(locs, grid, scale / 2.0), 0)
    meas_sum = np.sum(vals / 2 + 0.5)  # sum over the n columns of ones((G, n)) * val / 2 + 0.5: the same for every grid point
    meas_var = renormalize(np.mean(meas_map, 1))
    meas_var = meas_var * (scale - 50.0 * scale) + 50.0 * scale  # rescale [0, 1] -> [50 scale, scale]
    posterior_var = 1.0 / (1.0 / prior_var + n / meas_var)
    posterior = posterior_var * (prior / prior_var + meas_sum / meas_var)
    return posterior, posterior_var

"""Workspace <-> search-space helpers used by the controller constructor.

Pure index/affine bookkeeping on small numpy arrays (no hot-path arithmetic);
same call signatures as franka_test/scripts/franka/franka_utils.py:16-46 so the
callers' imports keep working.
"""
import numpy as np


def ws_conversion(pt, in_dim, out_dim):
    """Affine map of the leading len(in_dim) coordinates from in_dim to out_dim limits."""
    in_dim = np.atleast_2d(in_dim)
    out_dim = np.atleast_2d(out_dim)
    span_in = in_dim[:, 1] - in_dim[:, 0]
    span_out = out_dim[:, 1] - out_dim[:, 0]
    n = len(span_in)
    head = pt[:n] if np.ndim(pt) == 1 else pt[:, :n]
    return (head - in_dim[:, 0]) / span_in * span_out + out_dim[:, 0]


def find_non_vel_locs(states):
    """Split a state string into lower-case (position) and upper-case (velocity) slots."""
    pos = [(i, s) for i, s in enumerate(states) if s == s.lower()]
    non_vel_locs = np.array([i for i, _ in pos])
    vel_locs = [i for i, s in enumerate(states) if s == s.upper()]
    return non_vel_locs, vel_locs, "".join(s for _, s in pos)

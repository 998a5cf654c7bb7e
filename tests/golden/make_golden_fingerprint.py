#!/usr/bin/env python
"""Golden vectors of the fingerprint belief update, recorded from the LIVE reference
(/root/reference/franka_test/scripts/dist_modules/fingerprint_module.py: FingerprintDist.update_prior :539-589,
meas_footprint_vec :417-424, process_meas :470-478, build_grid :504-519).

The module imports matplotlib, the plotting package, numa, psutil ... none of which the belief arithmetic uses and
most of which this image lacks: they are replaced by empty stand-in modules before the import (the reference source is
imported in place, never copied).

    python tests/golden/make_golden_fingerprint.py      # rewrites tests/golden/fingerprint_*.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SCRIPTS = "/root/reference/franka_test/scripts"

SUB = 13  # grid stride of the recorded fields

CASES = {
    # name: states, workspace limits, measurements per update, updates, threshold / clip of process_meas
    "xy": dict(states="xy", lims=[[-1.0, 1.0], [-1.0, 1.0]], n=[7, 1, 12], thresh=None, clip=None),
    "xyw": dict(states="xyw", lims=[[-1.0, 1.0], [-1.0, 1.0], [-2.0, 2.0]], n=[5, 9], thresh=0.35, clip=1.2),
    "xyz_one": dict(states="xyz", lims=[[-1.0, 1.0], [-0.5, 0.8], [0.1, 0.6]], n=[1, 1, 3], thresh=0.2, clip=2.0),
}


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return m

    stub("termcolor", cprint=lambda *a, **k: None, colored=lambda s, *a, **k: s)
    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    stub("plotting")
    stub("plotting.plotting_matplotlib", EvalPlotter=object, set_mpl_format=lambda *a, **k: None)
    stub("numa")
    stub("psutil")
    if REF_SCRIPTS not in sys.path:
        sys.path.insert(0, REF_SCRIPTS)
    # under a private package name: this repo has a dist_modules package of its own (the device mirror)
    import importlib
    import importlib.machinery
    if "_ref_dist_modules" not in sys.modules:
        pkg = types.ModuleType("_ref_dist_modules")
        pkg.__path__ = [os.path.join(REF_SCRIPTS, "dist_modules")]
        pkg.__spec__ = importlib.machinery.ModuleSpec("_ref_dist_modules", None, is_package=True)
        pkg.__spec__.submodule_search_locations = pkg.__path__
        sys.modules["_ref_dist_modules"] = pkg
    return importlib.import_module("_ref_dist_modules.fingerprint_module")


def measurements(case, n, rng):
    lims = np.array(case["lims"], dtype=np.float64)
    locs = lims[:, 0] + rng.random((n, len(lims))) * (lims[:, 1] - lims[:, 0])
    vals = rng.random(n) * (1.5 if case["thresh"] is not None else 1.0)
    return locs, vals


def record(fm, name, case):
    rng = np.random.default_rng(17)
    fd = fm.FingerprintDist(explr_states=case["states"], plot_idx=[0, 1], capacity=64, lims=[list(x) for x in case["lims"]],
                            thresh=case["thresh"], clip=case["clip"], name=("a", "b", "c"))
    # the fixtures stay small: the 50^D grid is rebuilt by the tests from the limits (its strided rows are recorded
    # as a check), fields over the grid are recorded on every SUB-th point
    out = {"grid_rows": fd.grid[::SUB].copy(), "scale": np.array(fd.scale), "lims_scaled": fd.lims.copy(),
           "grid_points": np.array(fd.grid.shape[0])}
    for k, n in enumerate(case["n"]):
        locs, vals = measurements(case, n, rng)
        if n == 1:
            fd.push(locs[0], vals[0])
        else:
            fd.push_batch(locs, vals)
        out[f"u{k}/locs"], out[f"u{k}/vals"] = locs, vals
        out[f"u{k}/processed"] = fd.get_meas(separate=True)[1].copy()
        out[f"u{k}/meas_map"] = fm.meas_footprint_vec(samples=fd.grid, explr_idx=fd.update_idx, locs=locs, std=fd.scale / 2.)[::SUB]
        fd.update_prior()
        out[f"u{k}/prior"], out[f"u{k}/prior_var"] = fd.prior[::SUB].copy(), fd.prior_var[::SUB].copy()
        out[f"u{k}/prior_sum"], out[f"u{k}/prior_var_sum"] = np.array(fd.prior.sum()), np.array(fd.prior_var.sum())
    out["n_updates"] = np.array(len(case["n"]))
    np.savez_compressed(os.path.join(HERE, f"fingerprint_{name}.npz"), **out)
    print(name, "->", len(out), "arrays, grid", fd.grid.shape)


if __name__ == "__main__":
    fm = import_reference()
    for name, case in CASES.items():
        record(fm, name, case)

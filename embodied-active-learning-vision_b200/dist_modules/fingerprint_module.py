"""Per-fingerprint belief over the search grid, updated on the GPU.

Mirror of the reference's ``FingerprintDist`` (franka_test/scripts/dist_modules/fingerprint_module.py:426-627): same
constructor, buffers, ``push`` / ``push_batch`` / ``clear_batch`` / ``get_meas`` / ``process_meas`` / ``update_prior`` /
``pdf``.  The belief {prior, prior_var} lives in HBM as float64 (the reference keeps numpy float64 arrays) and
``update_prior`` runs klerg_belief_update (csrc/klerg_belief.cu) on it; ``prior`` / ``prior_var`` read back as numpy on
access.  ``pdf`` at arbitrary samples is the reference's scipy RBFInterpolator solve and stays on the host; ``pdf`` on
the grid (``use_grid=True``, the p_k the K-target planner consumes) never leaves the device when ``as_tensor=True``.

Not built (out of scope): the periodic gaussian_filter smoothing (``smooth=True`` raises), the unused
``use_mask`` branch, save_results / plotting."""
import numpy as np
import torch

from control_torch import _cabi as cabi


class FingerprintDist(object):
    def __init__(self, explr_states='xy', plot_idx=[0, 1], capacity=50000, scale=None, thresh=None, clip=None,
                 lims=[[-1, 1]] * 2, name=None, center=None, center_img=None, device=None):
        self.name = name
        self.explr_states = explr_states
        self.update_idx = np.arange(len(explr_states))
        self.plot_idx = plot_idx
        self.capacity = capacity
        self.scale = scale
        self.thresh = thresh
        self.clip = clip
        self.lims = np.array(lims)
        self.center = center
        self.center_img = center_img
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.init = False
        self.invert = False
        self.count = 0
        self._prior = None       # device float64 [G]
        self._prior_var = None
        self._scratch = None
        self.clear_batch()
        self.build_grid()

    # ---- grid (fingerprint_module.py:504-519) ----
    def build_grid(self):
        D = len(self.update_idx)
        self.extra_idx = tuple(i for i in range(D) if i not in self.plot_idx)
        if 'w' in self.explr_states:
            self.lims[self.explr_states.rfind('w')] *= 1.33
        self.lims *= 1.15
        num_samples = 50
        axes = np.linspace(*self.lims[self.update_idx].T, num_samples)          # [50, D]
        self.xy_mesh = np.meshgrid(*np.linspace(*self.lims[self.plot_idx].T, num_samples).T)
        self.xy_grid = np.stack([m.ravel() for m in self.xy_mesh], 1)
        self.mesh = np.meshgrid(*axes.T)
        self.grid = np.stack([m.ravel() for m in self.mesh], 1)
        self.num_samples = [num_samples] * D
        if self.scale is None:
            self.scale = np.max(axes[1] - axes[0]) * 2.5
        self._grid_dev = torch.from_numpy(np.ascontiguousarray(self.grid, dtype=np.float64)).to(self.device)

    # ---- belief, as numpy on access like the reference's attributes ----
    @property
    def prior(self):
        return None if self._prior is None else self._prior.cpu().numpy()

    @prior.setter
    def prior(self, v):
        self._prior = None if v is None else torch.as_tensor(np.asarray(v, dtype=np.float64)).to(self.device).contiguous()

    @property
    def prior_var(self):
        return None if self._prior_var is None else self._prior_var.cpu().numpy()

    @prior_var.setter
    def prior_var(self, v):
        self._prior_var = None if v is None else torch.as_tensor(np.asarray(v, dtype=np.float64)).to(self.device).contiguous()

    def init_uniform_grid(self, x):
        assert len(x.shape) > 1, 'Input needs to be a of size N x n'
        return np.full(x.shape[0], 0.5)

    # ---- measurements (fingerprint_module.py:470-502, 608-627) ----
    def process_meas(self, x):
        if self.thresh is None:
            return x
        t = self.thresh - x
        t[t > 0] /= self.thresh
        t[t < 0] /= (self.clip - self.thresh)
        return np.tanh(t)

    def get_meas(self, separate=False):
        if not (self.position > 0 or self.full_buffer):
            raise ValueError("need measurements to format")
        n = self.capacity if self.full_buffer else self.position
        locs = self.env_path[:n].copy()
        vals = self.process_meas(self.env_path_val[:n].copy())
        return (locs, vals) if separate else zip(locs, vals)

    def format_meas(self, scale):
        locs, val = self.get_meas(separate=True)
        return {'scale': scale, 'locs': locs, 'std': val}

    def push(self, state, val):
        if (not self.full_buffer) and (self.position + 1 == self.capacity):
            self.full_buffer = True
        self.env_path[self.position] = state
        self.env_path_val[self.position] = val
        self.position = (self.position + 1) % self.capacity

    def push_batch(self, state, val):
        k = val.shape[0]
        if (not self.full_buffer) and (self.position + k >= self.capacity):
            self.full_buffer = True
        self.env_path[self.position:self.position + k] = state
        self.env_path_val[self.position:self.position + k] = val
        self.position = (self.position + k) % self.capacity

    def clear_batch(self):
        self.full_buffer = False
        self.position = 0
        self.env_path = np.empty([self.capacity, len(self.explr_states)])
        self.env_path_val = np.empty(self.capacity)

    # ---- the belief update (fingerprint_module.py:539-589) ----
    def update_prior(self, debug_plots=False, smooth=False):
        if smooth:
            raise NotImplementedError("FingerprintDist.update_prior(smooth=True): the gaussian_filter pass is not built")
        G = self.grid.shape[0]
        if self._prior is None:
            self._prior = torch.full((G,), 0.5, dtype=torch.float64, device=self.device)
            self._prior_var = torch.full((G,), 2.0, dtype=torch.float64, device=self.device)
        loc, val = self.get_meas(separate=True)
        if loc.ndim < 2:
            loc = loc[None]
        n = loc.shape[0]
        lib = cabi.load()
        need = int(lib.klerg_belief_scratch_bytes(G, n))
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        locs_dev = torch.from_numpy(np.ascontiguousarray(loc[:, self.update_idx], dtype=np.float64)).to(self.device)
        meas_sum = float(np.sum(np.asarray(val, dtype=np.float64) / 2 + 0.5))
        with torch.cuda.device(self.device):
            cabi.check(lib.klerg_belief_update(
                cabi.ptr(self._grid_dev), G, self.grid.shape[1], cabi.ptr(locs_dev), n, float(self.scale), meas_sum,
                cabi.ptr(self._prior), cabi.ptr(self._prior_var), cabi.ptr(self._prior), cabi.ptr(self._prior_var),
                cabi.ptr(self._scratch), cabi.stream_ptr()), "klerg_belief_update")
        self.count += n
        self.clear_batch()

    # ---- density (fingerprint_module.py:591-606) ----
    def pdf(self, samples, override_invert=False, plot=False, use_grid=False, as_tensor=False):
        if use_grid:
            samples = self.grid
        if not (self.init and self._prior is not None):
            vals = self.init_uniform_grid(samples)
            return torch.from_numpy(vals).to(self.device) if as_tensor else vals
        if use_grid:
            dist = self._prior.clone()
        else:
            from scipy.interpolate import RBFInterpolator
            dist = torch.from_numpy(RBFInterpolator(self.grid, self.prior, kernel='linear')(samples)).to(self.device)
        if self.invert and not override_invert:
            dist = -dist + dist.max() + dist.min()
        return dist if as_tensor else dist.cpu().numpy()

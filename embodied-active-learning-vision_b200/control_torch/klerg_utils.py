"""Pairwise-kernel utilities with the reference's names and argument meaning
(franka_test/scripts/control_torch/klerg_utils.py), executed on the GPU.

Inputs may be CPU or CUDA tensors; the result lives where ``samples`` / ``dist``
lives.  CPU inputs are copied to the device, the CUDA kernel runs, and the
result is copied back - there is no host implementation.
"""
import torch

from . import _cabi as cabi
from . import engine


def _f32_dev(t):
    cabi.require_cuda()
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.detach().to(device="cuda", dtype=torch.float32).contiguous()


def _spec(traj_cols, explr_idx, std, nu):
    explr = [int(i) for i in torch.as_tensor(explr_idx).reshape(-1).tolist()]
    scale = torch.as_tensor(std, dtype=torch.float32).reshape(-1).tolist()
    if len(scale) == 1 and len(explr) > 1:
        scale = scale * len(explr)
    nu = float(torch.as_tensor(nu).reshape(-1)[0])
    return cabi.kernel_spec(len(explr), traj_cols, explr, scale, nu)


def _pairwise(mode, traj, samples, explr_idx, std, nu):
    spec = _spec(traj.shape[1], explr_idx, std, nu)
    s_dev = _f32_dev(samples)
    packed = engine.pack_samples(spec, s_dev)
    out, _ = engine.footprint(spec, mode, _f32_dev(traj), packed, s_dev.shape[0])
    return out[0, : s_dev.shape[0]].to(device=samples.device, dtype=samples.dtype)


def _psi(traj, samples, std, nu, want_psi, want_dpsi):
    import ctypes as C
    t = _f32_dev(traj).reshape(-1, traj.shape[-1])
    s = _f32_dev(samples).reshape(-1, samples.shape[-1])
    D = s.shape[1]
    spec = _spec(D, list(range(D)), std, nu)
    n, T = s.shape[0], t.shape[0]
    psi = torch.empty((n, T), dtype=torch.float32, device=s.device) if want_psi else None
    dpsi = torch.empty((n, T, D), dtype=torch.float32, device=s.device) if want_dpsi else None
    cabi.check(cabi.load().klerg_psi_matrix(C.byref(spec), cabi.ptr(t), T, cabi.ptr(s), n, cabi.ptr(psi), cabi.ptr(dpsi),
                                            cabi.stream_ptr()), "klerg_psi_matrix")
    return psi, dpsi


def psi_fn(traj, samples, std, nu):
    """psi[i, j] = exp(-0.5 sum_d (traj_jd - samples_id)^2 / std_d) / nu  (reference klerg_utils.py:7-10).

    traj [1, T, D] (or [T, D]), samples [N, 1, D] (or [N, D]) -> [N, T].  |std| is used, as every caller of the
    reference passes abs(std)."""
    psi, _ = _psi(traj, samples, std, nu, True, False)
    return psi.to(device=samples.device, dtype=samples.dtype)


def dpsi_dx_fn(x_explr, samples, std, nu):
    """-(x - s_i)/std * psi(x, s_i) for one state x [D] -> [N, D]  (reference klerg_utils.py:12-15)."""
    _, dpsi = _psi(x_explr.reshape(1, -1), samples, std, nu, False, True)
    return dpsi[:, 0, :].to(device=samples.device, dtype=samples.dtype)


def traj_footprint_vec(traj, samples, explr_idx, std, nu):
    """q_i = sum_j psi(traj_j[explr], samples_i)   (reference klerg_utils.py:17-22)."""
    return _pairwise(0, traj, samples, explr_idx, std, nu)


def traj_spread_vec(traj, samples, explr_idx, std, nu):
    """max_j psi(traj_j[explr], samples_i)         (reference klerg_utils.py:24-29)."""
    return _pairwise(1, traj, samples, explr_idx, std, nu)


def kldiv_grad_vec(x, samples, explr_idx, std, importance_ratio, nu):
    """sum_i ratio_i * d psi(x, s_i)/dx, zeros outside explr_idx (reference :12-15,31-36)."""
    spec = _spec(x.shape[0], explr_idx, std, nu)
    s_dev = _f32_dev(samples)
    packed = engine.pack_samples(spec, s_dev)
    g = engine.kl_gradient(spec, _f32_dev(x).unsqueeze(0), packed, s_dev.shape[0], _f32_dev(importance_ratio))
    return g[0].to(device=x.device, dtype=x.dtype)


def cost_norm(dist):
    """In place: NaN -> 1e-6, then dist /= dist.sum()  (reference :38-42)."""
    d = _f32_dev(dist).clone()
    engine.cost_norm_(d)
    dist.copy_(d.to(device=dist.device, dtype=dist.dtype))
    return dist


def renormalize(dist, dim=None, min_val=1e-6):
    """x/sum -> clamp(min_val) -> divide by the max (closed form of reference :45-58)."""
    if dim is None:
        flat = _f32_dev(dist).reshape(-1)
        return engine.renormalize(flat, min_val).reshape(dist.shape).to(device=dist.device, dtype=dist.dtype)
    moved = _f32_dev(dist).movedim(dim, -1).contiguous()
    rows = moved.reshape(-1, moved.shape[-1])
    out = torch.stack([engine.renormalize(r.contiguous(), min_val) for r in rows]).reshape(moved.shape)
    return out.movedim(-1, dim).to(device=dist.device, dtype=dist.dtype)


class Lambda(torch.nn.Module):
    """Wrap ``func(x, *vars)`` as a module (reference :60-69)."""

    def __init__(self, func, vars):
        super().__init__()
        self.func = func
        self.vars = vars

    def forward(self, x):
        return self.func(x, *self.vars)

    def extra_repr(self):
        return f"{self.func.__name__}(x, *vars) with {len(self.vars)} vars"

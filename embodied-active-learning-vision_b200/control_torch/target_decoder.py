"""Device evaluation of the VAE sensor model's target density.

The reference hands its CVAE straight to the controller as ``target_dist``
(dist_modules/sensor_main_module.py:99-102) and ``Robot.get_target_dist`` calls
``target_dist.pdf_torch(samples)`` (control_torch/klerg.py:463), i.e.
``VAE.pdf_torch`` (vae/vae.py:244-275): the decoder MLP over ``[z || sample]``,
column(s) ``[:ylogvar_dim]`` -> clamp -> (mean over the z buffer) -> exp -> amax.

``DeviceTarget`` wraps such a model (duck-typed: ``decode`` = Linear ReLU Linear
ReLU Linear, ``z_samples``, ``ylogvar_dim``, ``logvar_lims``, ``init``, optional
``dx`` / ``seed_x`` / ``use_buffer`` / ``z_buff``) and keeps the ``target_dist``
duck type (``pdf_torch``, ``init_uniform_grid``, ``device``, ``dtype``), so the
controller - or any other caller of ``pdf_torch`` - uses it unchanged.  The
weights are re-read from the wrapped model on every call (the trainer keeps
updating it, sensor_main_module.py:311-339) and the arithmetic runs in
``klerg_target_decoder_pdf`` (tcgen05 tensor cores, 3xTF32).  There is no CPU
fallback: a model whose decoder does not have the supported shape raises
``NotImplementedError``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi as cabi


def _linears(decode):
    """[Linear, Linear, Linear] of a Linear-ReLU-Linear-ReLU-Linear decoder (vae.py:77-84) or None."""
    try:
        mods = list(decode)
    except TypeError:
        return None
    lin = [m for m in mods if isinstance(m, torch.nn.Linear)]
    act = [m for m in mods if not isinstance(m, torch.nn.Linear)]
    if len(lin) != 3 or len(mods) != 5 or not all(isinstance(m, torch.nn.ReLU) for m in act):
        return None
    if not (isinstance(mods[0], torch.nn.Linear) and isinstance(mods[2], torch.nn.Linear)
            and isinstance(mods[4], torch.nn.Linear)):
        return None
    return lin


def is_decoder_model(target_dist):
    """True for objects shaped like the reference's VAE (vae/vae.py:11-110)."""
    return all(hasattr(target_dist, n) for n in ("decode", "z_samples", "ylogvar_dim", "logvar_lims", "init"))


def decoder_supported(model):
    """True if the model's decoder is one the kernel is built for (Linear-ReLU-Linear-ReLU-Linear with widths the
    packed layout accepts).  The controller evaluates any other VAE through the model's own ``pdf_torch``."""
    if not is_decoder_model(model):
        return False
    lin = _linears(model.decode)
    if lin is None or getattr(model, "use_chunk_decode", False):
        return False
    l1, l2, l3 = lin
    try:
        zd = int(model.z_samples.shape[-1])
        nz = int(model.z_buff.get_samples().shape[0]) if getattr(model, "use_buffer", False) else 1
        sd, nl = l1.in_features - zd, int(model.ylogvar_dim)
    except Exception:  # noqa: BLE001 - an object that only looks like the VAE
        return False
    if l2.in_features != l1.out_features or l3.in_features != l2.out_features or l3.out_features < nl or sd < 1:
        return False
    return cabi.load().klerg_target_decoder_packed_bytes(sd, zd, max(nz, 1), l1.out_features, l2.out_features, nl) != 0


class DeviceTarget:
    """``target_dist`` whose ``pdf_torch`` runs on the GPU (see module docstring)."""

    def __init__(self, model, device=None):
        cabi.require_cuda()
        if not is_decoder_model(model):
            raise TypeError("DeviceTarget wraps a VAE-like model (decode, z_samples, ylogvar_dim, logvar_lims, init)")
        self.model = model
        self.cuda = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = str(self.cuda)
        self.dtype = torch.float32
        self._packed = None
        self._fault = torch.zeros(1, dtype=torch.int32, device=self.cuda)
        self._dims = None
        self.stats = dict(packs=0, evals=0)

    # -- the wrapped model's own attributes stay reachable (callers poke at .init, .update_dist, ...)
    def __getattr__(self, name):
        return getattr(self.__dict__["model"], name)

    def init_uniform_grid(self, x):
        """vae.py:230-236: ``x.sum(1)**0`` (ones, NaN-free) on the device the samples live on."""
        return torch.ones(x.shape[0], dtype=torch.float32, device=x.device)

    # ------------------------------------------------------------------ weights
    def _z_rows(self):
        m = self.model
        if getattr(m, "use_buffer", False):
            return m.z_buff.get_samples()  # [z_mem, z_dim] (vae.py:253-257)
        return m.z_samples  # [1, z_dim] (vae.py:258-259)

    def refresh(self):
        """Copy the decoder weights and z to the device and pack them for the kernel."""
        lib = cabi.load()
        m = self.model
        lin = _linears(m.decode)
        if lin is None:
            raise NotImplementedError("target decoder: only Linear-ReLU-Linear-ReLU-Linear decoders are built (vae.py:77-84)")
        l1, l2, l3 = lin
        z = self._z_rows().detach().to(torch.float32).reshape(-1, int(m.z_samples.shape[-1]))
        zd, nz = int(z.shape[1]), int(z.shape[0])
        sd = l1.in_features - zd
        nl = int(m.ylogvar_dim)
        h1, h2 = l1.out_features, l2.out_features
        if l2.in_features != h1 or l3.in_features != h2 or l3.out_features < nl or sd < 1:
            raise NotImplementedError("target decoder: inconsistent layer shapes")
        dims = (sd, zd, nz, h1, h2, nl)
        parts = [l1.weight, l1.bias, z, l2.weight, l2.bias, l3.weight[:nl], l3.bias[:nl]]
        # Unchanged weights are not packed again: a tensor's version counter moves with every in-place update (an
        # optimiser step), its storage pointer with every replacement.
        key = (dims,) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in parts)
        if self._packed is not None and key == getattr(self, "_pack_key", None):
            return
        nbytes = lib.klerg_target_decoder_packed_bytes(*dims)
        if nbytes == 0:
            cabi.check(-1, "klerg_target_decoder_packed_bytes")
        if self._packed is None or self._packed.numel() < nbytes + 128:
            self._packed = torch.empty(nbytes + 128, dtype=torch.uint8, device=self.cuda)
        base = self._packed.data_ptr()
        self._packed_ptr = (base + 127) // 128 * 128
        # one flat H2D copy of everything the pack kernel reads
        flat = torch.cat([p.detach().to(torch.float32).reshape(-1).cpu() for p in parts])
        dev = flat.to(self.cuda, non_blocking=True)
        offs, o = [], 0
        for p in parts:
            offs.append(o)
            o += p.numel()
        ptrs = [C.c_void_p(dev.data_ptr() + 4 * o) for o in offs]
        cabi.check(lib.klerg_target_decoder_pack(*ptrs, *dims, C.c_void_p(self._packed_ptr), cabi.stream_ptr()),
                   "klerg_target_decoder_pack")
        self._staging = dev  # keep alive until the stream has consumed it
        self._dims = dims
        self._pack_key = key
        self.stats["packs"] += 1

    # ------------------------------------------------------------------ density
    @torch.no_grad()
    def pdf_torch(self, samples):
        """Same contract as VAE.pdf_torch: samples [N, s_dim] (host or device) -> p [N] on ``self.device``."""
        m = self.model
        samples = samples.to(device=self.cuda, dtype=torch.float32)
        if not bool(m.init):
            return self.init_uniform_grid(samples)
        self.refresh()
        return self.evaluate_packed(samples)

    @torch.no_grad()
    def evaluate_packed(self, samples):
        """The decoder kernel alone, on the weights of the last ``refresh()`` (samples [N, s_dim] on the device)."""
        m = self.model
        sd, zd, nz, h1, h2, nl = self._dims
        if samples.dim() != 2 or samples.shape[1] != sd:
            raise ValueError(f"target decoder: samples must be [N, {sd}]")
        samples = samples.contiguous()
        shift = None
        if getattr(m, "dx", False):
            shift = m.seed_x.detach().to(device=self.cuda, dtype=torch.float32).reshape(-1).contiguous()
        lo, hi = m.logvar_lims
        out = torch.empty(samples.shape[0], dtype=torch.float32, device=self.cuda)
        cabi.check(cabi.load().klerg_target_decoder_pdf(
            C.c_void_p(self._packed_ptr), sd, zd, nz, h1, h2, nl, cabi.ptr(samples), samples.shape[0],
            cabi.ptr(shift), float(lo), float(hi), cabi.ptr(out), cabi.ptr(self._fault), cabi.stream_ptr()),
            "klerg_target_decoder_pdf")
        self.stats["evals"] += 1
        return out.squeeze()

    def pdf(self, samples):
        """vae.py:238-242."""
        return self.pdf_torch(torch.as_tensor(samples)).cpu().numpy()

    def check_fault(self):
        """Raise if an in-kernel wait of any earlier ``pdf_torch`` timed out (one small D2H read)."""
        if int(self._fault.item()) != 0:
            raise RuntimeError("klerg_target_decoder_pdf: an in-kernel barrier wait timed out")

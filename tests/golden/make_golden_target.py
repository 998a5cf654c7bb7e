#!/usr/bin/env python
"""Golden vectors for the VAE target density, recorded from the LIVE reference VAE
(read-only /root/reference/franka_test/scripts/vae/vae.py):

    python tests/golden/make_golden_target.py        # rewrites tests/golden/target_*.npz

The reference model is built as the experiment configs build it (config/test_config.yaml:70-82),
its decoder is loaded with the seeded weights of cases.decoder_weights (so that the fixture stays
small: the weights are regenerated, not stored), ``update_dist`` produces z from the model's own
encoder, and ``pdf_torch`` is recorded.  Nothing here is imported by the product.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SCRIPTS = "/root/reference/franka_test/scripts"
sys.path.insert(0, HERE)
from cases import TARGET_CASES, decoder_weights, target_samples  # noqa: E402


def import_reference_vae():
    shim = types.ModuleType("termcolor")
    shim.cprint = lambda *a, **k: None
    shim.colored = lambda s, *a, **k: s
    sys.modules.setdefault("termcolor", shim)
    np.product = np.prod  # removed in numpy 2 (vae/vae_utils.py:21,30)
    if REF_SCRIPTS not in sys.path:
        sys.path.insert(0, REF_SCRIPTS)
    from vae.vae import VAE
    return VAE


def build_reference(VAE, case):
    torch.manual_seed(4)
    img = [3, 64, 64]
    nl = case["nl"]
    y_logvar_dim = nl if nl in (1, 3) else [3, 3]
    vae = VAE(img_dim=img, z_dim=case["zd"], s_dim=case["sd"], hidden_dim=list(case["hidden"]), y_logvar_dim=y_logvar_dim,
              CNNdict={'kernel_size': [3, 3, 5], 'stride': [2, 2, 3], 'channels': [10, 10, 20]}, dx=case["dx"])
    vae.eval()
    ws = decoder_weights(case, out_extra=vae.decode[-1].out_features - nl)
    with torch.no_grad():
        for lin, (w, b) in zip([m for m in vae.decode if isinstance(m, torch.nn.Linear)], ws):
            assert lin.weight.shape == w.shape, (lin.weight.shape, w.shape)
            lin.weight.copy_(w)
            lin.bias.copy_(b)
    if case["zbuf"]:
        vae.build_z_buffer(z_mem=case["zbuf"])
    return vae, ws


def main():
    VAE = import_reference_vae()
    for name, case in TARGET_CASES.items():
        vae, ws = build_reference(VAE, case)
        samples = target_samples(case)
        with torch.no_grad():
            p_uninit = vae.pdf_torch(samples.clone())
        g = torch.Generator().manual_seed(9)
        for _ in range(max(1, case["zbuf"] + 1)):  # one more push than rows: the ring wraps (vae_buffer.py:115-120)
            xr = torch.rand(1, case["sd"], generator=g) * 2 - 1
            y = torch.rand(1, 3, 64, 64, generator=g)
            vae.update_dist(xr, y)
        z_rows = vae.z_buff.get_samples() if case["zbuf"] else vae.z_samples.clone()
        with torch.no_grad():
            p = vae.pdf_torch(samples.clone())
        out = dict(z_rows=z_rows.numpy(), seed_x=vae.seed_x.numpy(), p=p.numpy(), p_uninit=p_uninit.numpy(),
                   out_features=np.array(vae.decode[-1].out_features))
        np.savez_compressed(os.path.join(HERE, f"target_{name}.npz"), **out)
        print(name, "p range", float(p.min()), float(p.max()), "clamped lo/hi:",
              int((p <= np.exp(-10.0) * 1.0000001).sum()), int((p >= np.exp(2.0) * 0.9999999).sum()), "z rows", tuple(z_rows.shape))


if __name__ == "__main__":
    main()

"""CPU oracle for the VAE target density.  TEST INFRASTRUCTURE ONLY.

torch-CPU fp32 restatement of ``VAE.pdf_torch`` of the reference
(``franka_test/scripts/vae/vae.py:244-275``) and of the pieces of the model it
reads (decoder ``vae.py:77-84``, ``init_uniform_grid`` ``vae.py:230-236``, z
buffer ``vae/vae_buffer.py:87-136``).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU legs may import it; the product package never does.

Pinning: ``tests/golden/make_golden_target.py`` imports the live reference VAE
from ``/root/reference`` (with ``np.product = np.prod`` for numpy 2 and the
``termcolor`` stand-in), loads seeded decoder weights, runs ``update_dist`` and
records ``pdf_torch`` outputs into ``tests/golden/target_*.npz``;
``tests/test_target_oracle.py`` checks this module against those vectors.
"""
from __future__ import annotations

import torch


def decode_mlp(latent, weights):
    """Linear ReLU Linear ReLU Linear (vae.py:77-84); weights = [(W, b)] * 3 in torch.nn.Linear layout."""
    h = latent
    for i, (w, b) in enumerate(weights):
        h = torch.nn.functional.linear(h, w, b)
        if i + 1 < len(weights):
            h = torch.relu(h)
    return h


@torch.no_grad()
def vae_pdf(samples, weights, z_rows, n_logvar, logvar_lims=(-10, 2), shift=None, initialised=True):
    """p(s) of VAE.pdf_torch (vae.py:244-275).

    samples [N, s_dim]; z_rows [n_z, z_dim] (one row = ``z_samples``, several = the z buffer); ``shift`` =
    ``seed_x`` for dx models (vae.py:249-250)."""
    samples = samples.to(torch.float32)
    if not initialised:
        return samples.sum(1) ** 0  # init_uniform_grid (vae.py:235)
    if shift is not None:
        samples = samples - shift
    n = samples.shape[0]
    latent = torch.vstack([torch.cat([zs.repeat(n, 1), samples], dim=1) for zs in z_rows])  # vae.py:256 / 259-260
    y_out = decode_mlp(latent, weights)
    var_data = torch.clamp(y_out[:, :n_logvar], *logvar_lims)  # vae.py:265-266
    if z_rows.shape[0] > 1:
        var_data = torch.mean(var_data.reshape(z_rows.shape[0], n, n_logvar), 0)  # vae.py:268-270
    var_data = torch.exp(var_data)
    return torch.amax(var_data, 1).squeeze()  # vae.py:273-275


@torch.no_grad()
def trainer_spread_grade(traj, samples, std, pdf_torch, xi=4.0):
    """The trainer's per-iteration "spread" and "grade" (dist_modules/trainer_module.py:511-538), with
    ``traj_spread_vec`` restated from control_torch/klerg_utils.py:24-29 and ``pdf_torch`` = the model's density
    (vae.py:244-275).  traj [M, D] = all replay-buffer states, samples [N, D].  Returns
    (spread, grade, max_q, entropy_dist)."""
    dim = samples.shape[1]
    std_t = torch.tensor([std] * dim)
    inner = torch.square(traj[:, :dim].unsqueeze(0) - samples.unsqueeze(1)) / torch.abs(std_t)  # klerg_utils.py:7-10
    max_q = torch.amax(torch.exp(-0.5 * torch.sum(inner, 2)) / 1.0, 1)  # traj_spread_vec, nu = 1 (trainer_module.py:518)
    max_q = max_q / torch.max(max_q)
    spread = max_q.mean()
    entropy_dist = pdf_torch(samples)
    entropy_dist = entropy_dist ** spread
    entropy_dist = entropy_dist / entropy_dist.max()
    grade = torch.clamp(10.0 ** (-torch.log10(entropy_dist.min()) - xi), max=0.01)
    return spread, grade, max_q, entropy_dist

"""Arithmetic of the tensor-core history pass (klerg_footprint_tc.cu), modelled in numpy: the exponent
e_ij = |sc_i|^2 + |xc_j|^2 - 2 xc_j . sc_i as ONE K = 8 bilinear form a_i . b_j, evaluated as the three tf32 products
a_lo b_hi + a_hi b_lo + a_hi b_hi with fp32 accumulation (3xTF32).  Checks the algebra (a . b == squared distance) and
the error bound the kernel's radius guard relies on: within |xc|^2 <= 100 the exponent is good to ~1e-4 absolute, i.e.
psi = 2^-e to < 1e-4 relative for every pair that matters."""
import numpy as np
import pytest


def tf32_rna(x):
    """cvt.rna.tf32.f32: round fp32 to the 10-bit mantissa the MMA reads (nearest, ties away from zero)."""
    bits = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    return ((bits + np.uint64(0x1000)) & np.uint64(0xFFFFE000)).astype(np.uint32).view(np.float32)


def tf32_trunc(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x, rounded=True):
    """hi + lo as the kernel forms them: both halves ROUNDED to tf32 (truncating them - a plain mask, and the hardware's
    own treatment of the low 13 bits - doubles each half's error and quadruples the dropped lo * lo term)."""
    cvt = tf32_rna if rounded else tf32_trunc
    hi = cvt(x)
    lo = cvt((x.astype(np.float32) - hi).astype(np.float32))
    return hi, lo


def mma_3xtf32(a, b, per_product_rounding=False):
    """D = a_lo b_hi + a_hi b_lo + a_hi b_hi, each K = 8 product exact (11 x 11 bits), fp32 accumulation."""
    a_hi, a_lo = split(a)
    b_hi, b_lo = split(b)
    acc = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    for x, y in ((a_lo, b_hi), (a_hi, b_lo), (a_hi, b_hi)):
        if per_product_rounding:  # pessimistic: the accumulator rounds after every one of the 24 products
            for k in range(a.shape[1]):
                acc = (acc + (x[:, k:k + 1].astype(np.float64) * y[None, :, k].astype(np.float64)).astype(np.float32)).astype(np.float32)
        else:  # one rounding per MMA: the K = 8 products are summed exactly before they meet the accumulator
            acc = (acc.astype(np.float64) + x.astype(np.float64) @ y.astype(np.float64).T).astype(np.float32)
    return acc


def operands(samples, states, centre):
    D = samples.shape[1]
    sc = (samples - centre).astype(np.float32)
    xc = (states - centre).astype(np.float32)
    a = np.zeros((samples.shape[0], 8), dtype=np.float32)
    b = np.zeros((states.shape[0], 8), dtype=np.float32)
    a[:, :D], a[:, 6], a[:, 7] = sc, (sc.astype(np.float64) ** 2).sum(1), 1.0
    b[:, :D], b[:, 6], b[:, 7] = -2.0 * xc, 1.0, (xc.astype(np.float64) ** 2).sum(1)
    return a, b


@pytest.mark.parametrize("D,radius", [(6, 3.0), (6, 9.9), (3, 9.9), (2, 5.0)])
def test_bilinear_form_is_the_squared_distance_with_fp32_grade_error(D, radius):
    g = np.random.default_rng(D * 100 + int(radius))
    n, m = 400, 300
    states = g.normal(size=(m, D))
    states *= radius * g.random((m, 1)) / np.linalg.norm(states, axis=1, keepdims=True)  # |xc| <= radius around 0
    centre = 0.5 * (states.min(0) + states.max(0))
    # samples that matter: within a few kernel widths of some state (e <~ 30), plus far ones
    near = states[g.integers(0, m, n // 2)] + g.normal(scale=1.5, size=(n // 2, D))
    far = g.uniform(-1.6 * radius, 1.6 * radius, size=(n - n // 2, D))
    samples = np.vstack([near, far])
    a, b = operands(samples.astype(np.float32), states.astype(np.float32), centre.astype(np.float32))
    e_tc = mma_3xtf32(a, b).astype(np.float64)
    e_pess = mma_3xtf32(a, b, per_product_rounding=True).astype(np.float64)
    e_ref = ((samples.astype(np.float32).astype(np.float64)[:, None, :] - states.astype(np.float32).astype(np.float64)[None, :, :]) ** 2).sum(2)
    err = np.abs(e_tc - e_ref)
    matters = e_ref < 40.0  # psi > 1e-12
    # 2^-e relative error = ln2 * |de|
    assert matters.sum() > 1000
    assert np.log(2.0) * err[matters].max() < 1e-4, err[matters].max()
    # even if the accumulator rounded after every product the bound would be missed by less than a factor of two at
    # the very edge of the radius (the GPU test at that edge measures what the hardware does)
    assert np.log(2.0) * np.abs(e_pess - e_ref)[matters].max() < 2e-4
    # and nothing blows up for the far pairs (they underflow to 0 either way)
    assert np.isfinite(e_tc).all() and (e_tc[~matters] > 30.0).all()


def test_padding_rows_never_win():
    """Rows beyond the state list carry b = (0, .., 1, 1e30): e = |sc|^2 + 1e30 -> psi = 0 and never the minimum."""
    a, _ = operands(np.array([[0.3, -0.2, 0.1]], dtype=np.float32), np.zeros((1, 3), dtype=np.float32), np.zeros(3, dtype=np.float32))
    b = np.zeros((1, 8), dtype=np.float32)
    b[0, 6], b[0, 7] = 1.0, 1e30
    e = mma_3xtf32(a, b)
    assert np.isfinite(e).all() and e[0, 0] > 1e29

// Packed-FP32 (f32x2) pair arithmetic of the Gaussian-like kernel psi (klerg_utils.py:7-10).
//
// On sm_100a FADD2/FMUL2/FFMA2 process two fp32 lanes per instruction.  Measured on B200
// (tools/microbench.cu): the FP32 pipe delivers ~127 lane-ops/clk/SM with the packed forms
// but only ~85 with three-register scalar FFMA (register-file bandwidth), and the packed
// forms halve the issue slots, which leaves room for the MUFU.EX2 (16 lanes/clk/SM) that
// bounds the forward pair.  All kernels therefore work on PAIRS OF SAMPLES (lo, hi) against
// one state whose coordinates are duplicated into both halves.
//
// Coordinates are pre-scaled (x' = x * sqrt(0.5*log2e/|scale_d|)) so that
//   psi = 2^-(sum_d (x'_d - s'_d)^2)    -> D FADD + 1 FMUL + (D-1) FFMA + 1 MUFU.EX2 per pair.
#pragma once
#include "klerg_common.cuh"

namespace klerg {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// 2^-x (negation folded into the MUFU operand)
__device__ __forceinline__ float ex2_neg(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(-x));
  return y;
}

// u64 slots per duplicated state row {x_d, x_d} in shared memory (even, so rows are 16 B aligned)
template <int D>
struct Row2 {
  static constexpr int DP = (D + 1) & ~1;
};

template <int D>
__device__ __forceinline__ void load_state2(const u64* __restrict__ sh, int j, u64 (&x)[D]) {
  constexpr int DP = Row2<D>::DP;
  const ulonglong2* r = reinterpret_cast<const ulonglong2*>(sh + j * DP);
#pragma unroll
  for (int h = 0; h < DP / 2; ++h) {
    const ulonglong2 v = r[h];
    x[2 * h] = v.x;
    if (2 * h + 1 < D) x[2 * h + 1] = v.y;
  }
}

// squared scaled distance of one duplicated state to a packed pair of samples; df kept for the gradient
template <int D>
__device__ __forceinline__ u64 sqdist2(const u64 (&x)[D], const u64 (&s)[D], u64 (&df)[D]) {
  df[0] = sub2(x[0], s[0]);
  u64 e = mul2(df[0], df[0]);
#pragma unroll
  for (int d = 1; d < D; ++d) {
    df[d] = sub2(x[d], s[d]);
    e = fma2(df[d], df[d], e);
  }
  return e;
}

// Forward pass of P packed sample pairs against rows [j0, j1) of duplicated state rows in shared memory.
//   MODE 0: acc[q] += {psi_lo, psi_hi}           (traj_footprint_vec, klerg_utils.py:17-22)
//   MODE 1: emin[2q..] = min_j squared distance  (traj_spread_vec, :24-29; psi of the min = max psi)
//   MODE 2: both
template <int D, int P, int MODE>
__device__ __forceinline__ void pair_forward(const u64* __restrict__ sh_x2, int j0, int j1, const u64 (&s2)[D][P],
                                             u64 (&acc)[P], float (&emin)[2 * P]) {
#pragma unroll 8
  for (int j = j0; j < j1; ++j) {
    u64 x[D];
    load_state2<D>(sh_x2, j, x);
#pragma unroll
    for (int q = 0; q < P; ++q) {
      u64 sq[D], df[D];
#pragma unroll
      for (int d = 0; d < D; ++d) sq[d] = s2[d][q];
      const u64 e = sqdist2<D>(x, sq, df);
      float e0, e1;
      unpack2(e, e0, e1);
      if (MODE != 1) acc[q] = add2(acc[q], pack2(ex2_neg(e0), ex2_neg(e1)));
      if (MODE != 0) {
        emin[2 * q] = fminf(emin[2 * q], e0);
        emin[2 * q + 1] = fminf(emin[2 * q + 1], e1);
      }
    }
  }
}
template <int D, int P, int MODE>
__device__ __forceinline__ void pair_forward(const u64* __restrict__ sh_x2, int T, const u64 (&s2)[D][P], u64 (&acc)[P],
                                             float (&emin)[2 * P]) {
  pair_forward<D, P, MODE>(sh_x2, 0, T, s2, acc, emin);
}

// Gradient pair update: WT duplicated states (registers) against one packed sample pair with
// packed importance weights w2:  acc[k][d] += w * psi * (x'_kd - s'_d)   (klerg_utils.py:12-15,31-36;
// the -1/(a*|scale|*nu) factor is applied once at the end, KernelDev::gfac).
// `full` = false skips the last state (warp-uniform): warps of one CTA may own WT-1 or WT states.
template <int D, int WT>
__device__ __forceinline__ void pair_gradient(const u64 (&xs2)[WT][D], const u64 (&s2)[D], u64 w2, u64 (&acc)[WT][D],
                                              bool full = true) {
#pragma unroll
  for (int k = 0; k < WT; ++k) {
    if (k == WT - 1 && !full) break;
    u64 df[D];
    const u64 e = sqdist2<D>(xs2[k], s2, df);
    float e0, e1;
    unpack2(e, e0, e1);
    const u64 wp = mul2(w2, pack2(ex2_neg(e0), ex2_neg(e1)));
#pragma unroll
    for (int d = 0; d < D; ++d) acc[k][d] = fma2(wp, df[d], acc[k][d]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Expanded form of the squared distance, in coordinates centred on the state set (c = a state in its middle):
//   e = |s - x|^2 = |sc|^2 + |xc|^2 - 2 xc . sc,      sc = s' - c,  xc = x' - c
// The per-sample term |sc|^2 and the per-state terms -2 xc_d, |xc|^2 are computed once, so a pair costs
//   forward : 1 FADD + D FFMA (e) + 1 FADD (sum)               = D + 2  lane-ops + 1 MUFU.EX2  (difference form: 2D + 1)
//   gradient: 1 FADD + D FFMA (e) + 1 FMUL (w psi) + D FFMA + 1 FADD = 2D + 3 lane-ops + 1 MUFU (difference form: 3D + 1)
// and, for D = 6, the forward pair is no longer bound by the FP32 pipe alone (8 lane-ops per MUFU: both pipes full).
// The gradient accumulates A_d = sum w psi sc_d and W = sum w psi; the caller forms xc_d W - A_d at the very end
// (in double).  Rounding: e is a difference of terms of size (R + |s - x|)^2 with R = radius of the state set around
// c, so its absolute error is ~1e-7 (R + 4)^2 for the pairs that matter (e < 17); callers use this form only while
// max |xc|^2 <= X_FORM_MAX_R2 (relative error of psi below ~3e-5: measured 2e-5 at R = 11.5 in a numpy model, 3.3e-5 worst
// case over 1e12 pairs of the 1e7 x 1e5 history pass at R <= 12) and fall back to the
// difference form otherwise.
constexpr float X_FORM_MAX_R2 = 100.f;

// floats per state row {-2 xc_0 .. -2 xc_{D-1}, |xc|^2, 0 ..}: 16-byte multiples, read as warp-wide broadcasts
template <int D>
struct RowX {
  static constexpr int NF = (D + 1 + 3) & ~3;
};

template <int D>
__device__ __forceinline__ void load_state_x(const float* __restrict__ sh_rows, int j, u64 (&m2x)[D], u64& x2n) {
  constexpr int NF = RowX<D>::NF;
  const float4* rp = reinterpret_cast<const float4*>(sh_rows + (size_t)j * NF);
  float r[NF];
#pragma unroll
  for (int h = 0; h < NF / 4; ++h) {
    const float4 v = rp[h];
    r[4 * h] = v.x;
    r[4 * h + 1] = v.y;
    r[4 * h + 2] = v.z;
    r[4 * h + 3] = v.w;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) m2x[d] = pack2(r[d], r[d]);  // one register, broadcast to both lanes by the FP2 operand form
  x2n = pack2(r[D], r[D]);
}

// Forward pass of P packed (centred) sample pairs against rows [j0, j1) of expanded-form state rows.
//   MODE 0: acc += psi   MODE 1: emin = min e   MODE 2: both (sum and max over the same rows)
// RB rows are evaluated together with their FFMA chains interleaved in program order (dimension outermost): the
// chain of one pair is D + 1 dependent instructions, and the compiler otherwise runs the chains one after the other
// through the same registers, which leaves the FP32 pipe waiting on its own latency.
template <int D, int P, int MODE, int RB>
__device__ __forceinline__ void pair_forward_x_rows(const float* __restrict__ sh_rows, int j, const u64 (&sc)[D][P],
                                                    const u64 (&s2n)[P], u64 (&acc)[P], float (&emin)[2 * P]) {
  u64 m2x[RB][D], e[RB][P];
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    u64 x2n;
    load_state_x<D>(sh_rows, j + r, m2x[r], x2n);
#pragma unroll
    for (int q = 0; q < P; ++q) e[r][q] = add2(s2n[q], x2n);
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int q = 0; q < P; ++q) e[r][q] = fma2(m2x[r][d], sc[d][q], e[r][q]);
  float ps[RB][P][2];
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int q = 0; q < P; ++q) {
      float e0, e1;
      unpack2(e[r][q], e0, e1);
      if (MODE != 1) {
        ps[r][q][0] = ex2_neg(e0);
        ps[r][q][1] = ex2_neg(e1);
      }
      if (MODE != 0) {
        emin[2 * q] = fminf(emin[2 * q], e0);
        emin[2 * q + 1] = fminf(emin[2 * q + 1], e1);
      }
    }
  if (MODE != 1) {
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int q = 0; q < P; ++q) acc[q] = add2(acc[q], pack2(ps[r][q][0], ps[r][q][1]));
  }
}

template <int D, int P, int MODE>
__device__ __forceinline__ void pair_forward_x(const float* __restrict__ sh_rows, int j0, int j1, const u64 (&sc)[D][P],
                                               const u64 (&s2n)[P], u64 (&acc)[P], float (&emin)[2 * P]) {
  constexpr int RB = P == 1 ? 4 : 2;  // 4 interleaved chains either way
  int j = j0;
#pragma unroll 2
  for (; j + RB <= j1; j += RB) pair_forward_x_rows<D, P, MODE, RB>(sh_rows, j, sc, s2n, acc, emin);
  for (; j < j1; ++j) pair_forward_x_rows<D, P, MODE, 1>(sh_rows, j, sc, s2n, acc, emin);
}

// Gradient pair update, expanded form: NS states (registers: -2 xc, |xc|^2 as scalars) against one packed pair of
// centred samples with |sc|^2 and the importance weights:  A[k][d] += w psi sc_d,  W[k] += w psi.
// The NS chains are interleaved in program order (see pair_forward_x_rows).  NS <= WT: a warp that owns one state
// fewer than the array holds calls the NS = WT - 1 instance (warp-uniform branch in the caller).
template <int D, int WT, int NS>
__device__ __forceinline__ void pair_gradient_x(const float (&m2x)[WT][D], const float (&x2n)[WT], const u64 (&sc)[D], u64 s2n,
                                                u64 w2, u64 (&A)[WT][D], u64 (&W)[WT]) {
  u64 e[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) e[k] = add2(s2n, pack2(x2n[k], x2n[k]));
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int k = 0; k < NS; ++k) e[k] = fma2(pack2(m2x[k][d], m2x[k][d]), sc[d], e[k]);
  u64 wp[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    float e0, e1;
    unpack2(e[k], e0, e1);
    wp[k] = mul2(w2, pack2(ex2_neg(e0), ex2_neg(e1)));
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) {
#pragma unroll
    for (int d = 0; d < D; ++d) A[k][d] = fma2(wp[k], sc[d], A[k][d]);
    W[k] = add2(W[k], wp[k]);
  }
}

// The same with the importance weight folded into the exponent: the caller hands in s2n = |sc|^2 - log2(w) per
// sample (+inf for w = 0), so 2^-e IS w psi: one multiply and one 64-bit shared load less per pair.
template <int D, int WT, int NS>
__device__ __forceinline__ void pair_gradient_xw(const float (&m2x)[WT][D], const float (&x2n)[WT], const u64 (&sc)[D], u64 s2n,
                                                 u64 (&A)[WT][D], u64 (&W)[WT]) {
  u64 e[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) e[k] = add2(s2n, pack2(x2n[k], x2n[k]));
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int k = 0; k < NS; ++k) e[k] = fma2(pack2(m2x[k][d], m2x[k][d]), sc[d], e[k]);
  u64 wp[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    float e0, e1;
    unpack2(e[k], e0, e1);
    wp[k] = pack2(ex2_neg(e0), ex2_neg(e1));
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) {
#pragma unroll
    for (int d = 0; d < D; ++d) A[k][d] = fma2(wp[k], sc[d], A[k][d]);
    W[k] = add2(W[k], wp[k]);
  }
}

}  // namespace klerg

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

// Forward pass of P packed sample pairs against T duplicated state rows in shared memory.
//   MODE 0: acc[q] += {psi_lo, psi_hi}           (traj_footprint_vec, klerg_utils.py:17-22)
//   MODE 1: emin[2q..] = min_j squared distance  (traj_spread_vec, :24-29; psi of the min = max psi)
template <int D, int P, int MODE>
__device__ __forceinline__ void pair_forward(const u64* __restrict__ sh_x2, int T, const u64 (&s2)[D][P], u64 (&acc)[P],
                                             float (&emin)[2 * P]) {
#pragma unroll 8
  for (int j = 0; j < T; ++j) {
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
      if (MODE == 0) {
        acc[q] = add2(acc[q], pack2(ex2_neg(e0), ex2_neg(e1)));
      } else {
        emin[2 * q] = fminf(emin[2 * q], e0);
        emin[2 * q + 1] = fminf(emin[2 * q + 1], e1);
      }
    }
  }
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

}  // namespace klerg

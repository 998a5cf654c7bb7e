// Shared device helpers for libklerg_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "klerg_b200.h"

namespace klerg {

// ---- workspace layout (see klerg_workspace_bytes) ---------------------------
// HEAD: [misc counters 64 x int][misc partials MAXBLK x 8 doubles]
//       [gradient partials GRAD_MAXBLK x MAX_H*MAX_D doubles][fused-eval region, see FUSED_*]
// then one SEG record per segment g: [counter + pad 64 B][MAXBLK x 2 doubles]
constexpr int MAXBLK = 9472;      // 148 SMs x 64
constexpr int GRAD_MAXBLK = 296;  // 148 SMs x 2
constexpr size_t HEAD_COUNTERS = 256;
constexpr size_t HEAD_MISC = (size_t)MAXBLK * 8 * sizeof(double);
constexpr size_t HEAD_GRAD = (size_t)GRAD_MAXBLK * KLERG_MAX_H * KLERG_MAX_D * sizeof(double);
// fused eval kernels (klerg_fused.cu): control words + this GPU's exchange mailbox (see "mailbox layout")
constexpr int FUSED_MAXG = 8;      // candidates per fused cost launch
constexpr size_t FUSED_CTRL = 256;  // u32: [5] sticky fault word; +64: debug stamps (18 x int64)

// ---- mailbox layout (symmetric across ranks; klerg_mailbox_bytes) --------------------------------------------
// All meeting points of the fused evals are flag-free "LL" exchanges: every value travels as 8-byte words
// {32 payload bits, 32-bit tag}; an aligned 8-byte store is single-copy atomic, so data and flag arrive together and
// a waiter simply polls the slot until the tag of the current exchange shows up.  No counters, nothing to reset.
// A double takes two words (16 B), an fp32 gradient partial one.  Slots are double-buffered by the parity of a
// device-resident launch counter (epoch) resp. exchange counter (xcount) kept in the mailbox header.
constexpr int MB_MAXW = 8;          // ranks
constexpr int LL_MAXBLK = 160;      // CTAs per rank that take part in an exchange (>= 148 SMs)
constexpr int LL_MAXNV = 2 * FUSED_MAXG;                      // doubles per CTA at a meeting
constexpr int LL_MAXHD = KLERG_MAX_H * (KLERG_MAX_D + 1);     // gather entries of one eval: H*D weighted sums + H weights
constexpr int LL_BUF_VALS = LL_MAXBLK * LL_MAXNV;             // polled values staged in shared memory (20 KB)
constexpr size_t MB_HDR = 256;      // u32 [0] epoch  [1] xcount  [2] xdone
constexpr size_t MB_X1 = (size_t)2 * MB_MAXW * LL_MAXBLK * 2 * 16;   // [par][rank][cta][2]   cross-rank one-hop meeting
constexpr size_t MB_L1 = (size_t)2 * LL_MAXBLK * LL_MAXNV * 16;      // [par][cta][nv]        local stage
constexpr size_t MB_R1 = (size_t)2 * MB_MAXW * LL_MAXNV * 16;        // [par][rank][nv]       rank stage
constexpr size_t MB_KL = (size_t)2 * LL_MAXBLK * LL_MAXNV * 16;      // [par][cta][nv]        KL partials -> finisher
constexpr size_t MB_GB = (size_t)2 * MB_MAXW * (LL_MAXHD + 16) * 16; // [par][rank][HD+16]    rank sums -> finisher
constexpr size_t MB_GP = (size_t)2 * LL_MAXHD * LL_MAXBLK * 8;       // [par][e][cta] {fp32, tag} CTA partials
constexpr size_t MB_OFF_X1 = MB_HDR;
constexpr size_t MB_OFF_L1 = MB_OFF_X1 + MB_X1;
constexpr size_t MB_OFF_R1 = MB_OFF_L1 + MB_L1;
constexpr size_t MB_OFF_KL = MB_OFF_R1 + MB_R1;
constexpr size_t MB_OFF_GB = MB_OFF_KL + MB_KL;
constexpr size_t MB_OFF_GP = MB_OFF_GB + MB_GB;
constexpr size_t MB_OFF_DBG = MB_OFF_GP + MB_GP;                     // [cta][8] globaltimer stamps (KLERG_STAMPS builds)
constexpr size_t MB_DBG = (size_t)LL_MAXBLK * 8 * 8;
// [par] rollout of the launch (states [H+1][S], rotations), published by its first CTA for the CTAs that start later;
// header word 4 + par holds epoch + 1 once it is complete
constexpr size_t MB_RO_FLOATS = (size_t)(KLERG_MAX_H + 1) * KLERG_MAX_S + (size_t)(2 * KLERG_MAX_H + 1) * 9;
constexpr size_t MB_OFF_RO = MB_OFF_DBG + MB_DBG;
constexpr size_t MB_RO = 2 * ((MB_RO_FLOATS * sizeof(float) + 15) & ~(size_t)15);
constexpr size_t MB_BYTES = MB_OFF_RO + MB_RO;

constexpr size_t FUSED_BYTES = FUSED_CTRL + MB_BYTES;  // single GPU: the mailbox lives inside the workspace
constexpr size_t HEAD_BYTES = HEAD_COUNTERS + HEAD_MISC + HEAD_GRAD + FUSED_BYTES;
constexpr size_t SEG_BYTES = 64 + (size_t)MAXBLK * 2 * sizeof(double);

__host__ __device__ inline int* ws_misc_counter(void* ws, int slot) { return (int*)ws + slot; }
__host__ __device__ inline double* ws_misc_partials(void* ws) { return (double*)((char*)ws + HEAD_COUNTERS); }
__host__ __device__ inline double* ws_grad_partials(void* ws) {
  return (double*)((char*)ws + HEAD_COUNTERS + HEAD_MISC);
}
__host__ __device__ inline char* ws_fused_base(void* ws) { return (char*)ws + HEAD_COUNTERS + HEAD_MISC + HEAD_GRAD; }
__host__ __device__ inline unsigned* ws_fused_ctrl(void* ws) { return (unsigned*)ws_fused_base(ws); }
__host__ __device__ inline void* ws_fused_mailbox(void* ws) { return ws_fused_base(ws) + FUSED_CTRL; }
__host__ __device__ inline int* ws_seg_counter(void* ws, int64_t g) {
  return (int*)((char*)ws + HEAD_BYTES + (size_t)g * SEG_BYTES);
}
__host__ __device__ inline double* ws_seg_partials(void* ws, int64_t g) {
  return (double*)((char*)ws + HEAD_BYTES + (size_t)g * SEG_BYTES + 64);
}

// ---- math -------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Loop-invariant kernel parameters of the serial recurrences (rollout, adjoint) are passed through shared
// memory once: ptxas treats constant-bank loads as free and otherwise re-issues LDC/LDCU inside those
// latency-critical loops, where each one sits on the dependency chain (measured: 140 cycles per rollout
// step instead of ~20).  A value loaded from (volatile) shared memory stays in its register.
struct PinnedParams {
  int S, A, H, kind;
  float dt;
};
__device__ __forceinline__ PinnedParams pin_params(float* s_scratch8, int S, int A, int H, int kind, float dt) {
  if (threadIdx.x == 0) {
    s_scratch8[0] = __int_as_float(S);
    s_scratch8[1] = __int_as_float(A);
    s_scratch8[2] = __int_as_float(H);
    s_scratch8[3] = __int_as_float(kind);
    s_scratch8[4] = dt;
  }
  __syncthreads();
  const volatile float* v = s_scratch8;
  PinnedParams p;
  p.S = __float_as_int(v[0]);
  p.A = __float_as_int(v[1]);
  p.H = __float_as_int(v[2]);
  p.kind = __float_as_int(v[3]);
  p.dt = v[4];
  return p;
}

// 0.5*log2(e): psi = exp(-0.5*sum d^2/scale) = 2^-(sum (d*a)^2), a = sqrt(HALF_LOG2E/scale)
constexpr double HALF_LOG2E = 0.72134752044448170368;

enum { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };

__device__ __forceinline__ double red_combine(int kind, double a, double b) {
  return kind == RED_SUM ? a + b : (kind == RED_MAX ? fmax(a, b) : fmin(a, b));
}
__device__ __forceinline__ double red_identity(int kind) {
  return kind == RED_SUM ? 0.0 : (kind == RED_MAX ? -INFINITY : INFINITY);
}

__device__ __forceinline__ double warp_reduce(int kind, double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_combine(kind, v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-level reduction of NQ quantities followed by a deterministic grid-level
// reduction done by the last block to finish ("ticket" pattern).  `nblk` blocks
// take part; block `blk` writes partials[blk*NQ + q]; the last block combines all
// partials in block order and writes out[q].  The counter is left at zero.
// Returns true in the last block (all threads).  Requires blockDim.x <= 1024.
template <int NQ>
__device__ __forceinline__ bool grid_reduce(const int (&kind)[NQ], double (&val)[NQ], double* partials,
                                            int* counter, int blk, int nblk, double* out) {
  __shared__ double sh_red[32][NQ];
  __shared__ int sh_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double v = warp_reduce(kind[q], val[q]);
    if (lane == 0) sh_red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      double v = sh_red[0][q];
      for (int w = 1; w < nwarp; ++w) v = red_combine(kind[q], v, sh_red[w][q]);
      partials[(size_t)blk * NQ + q] = v;
    }
    __threadfence();
    int ticket = atomicAdd(counter, 1);
    sh_last = (ticket == nblk - 1);
  }
  __syncthreads();
  const bool last = sh_last != 0;
  if (last) {
    __threadfence();
    // fixed-order combine: thread q-stripes, then a warp tree (order fixed by layout)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      double v = red_identity(kind[q]);
      for (int b = threadIdx.x; b < nblk; b += blockDim.x)
        v = red_combine(kind[q], v, __ldcg(&partials[(size_t)b * NQ + q]));
      v = warp_reduce(kind[q], v);
      __syncthreads();
      if (lane == 0) sh_red[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        double v = sh_red[0][q];
        for (int w = 1; w < nwarp; ++w) v = red_combine(kind[q], v, sh_red[w][q]);
        out[q] = v;
      }
      *counter = 0;
    }
  }
  return last;
}

// Sum / max of `world` rank blocks of [G][2] totals for segment g.
__device__ __forceinline__ void gather_totals(const double* totals, int world, int64_t G, int64_t g,
                                              double& sum, double& mx) {
  sum = 0.0;
  mx = -INFINITY;
  for (int r = 0; r < world; ++r) {
    sum += totals[((size_t)r * G + g) * 2 + 0];
    mx = fmax(mx, totals[((size_t)r * G + g) * 2 + 1]);
  }
}

struct KernelDev {
  int D, S;
  int explr[KLERG_MAX_D];
  float a[KLERG_MAX_D];       // sqrt(HALF_LOG2E/|scale|)
  float gfac[KLERG_MAX_D];    // -1/(a*|scale|*nu): scaled-difference sum -> dgdx
  float inv_nu;
  float x_r2;                 // largest |xc|^2 for which the expanded pair form is used (< 0: never)
};
extern int g_exact_pairs;     // KLERG_OPT_EXACT_PAIRS
extern int g_saturate_milli;  // KLERG_OPT_SATURATE_MILLI

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();
bool make_kernel_dev(const klerg_kernel_spec* k, KernelDev& out);

}  // namespace klerg

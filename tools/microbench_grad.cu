// Inner-loop probes of the gradient / forward pair passes in the kernel's own configuration (sm_100a):
// one CTA per SM, the warp's states in registers, sample tiles in shared memory, the pair functions of
// csrc/klerg_pair.cuh.  Variants isolate what bounds the loop: FP32-pipe work, the scalar-broadcast operand
// form of the packed instructions, shared-memory loads, dependent-chain latency.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -I embodied-active-learning-vision_b200/csrc \
//        -o tools/microbench_grad tools/microbench_grad.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "klerg_pair.cuh"

using namespace klerg;
namespace klerg { int g_exact_pairs = 0; }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int TS = 2048;  // samples per tile row

// rate probes for packed ops with a scalar-broadcast operand
template <int KIND>
__global__ void __launch_bounds__(256) probe(int iters, const float* __restrict__ in, float* out) {
  constexpr int NA = 8;
  float b[NA];
  u64 A2[NA], C2[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    b[i] = in[threadIdx.x + i];
    A2[i] = pack2(in[64 + threadIdx.x + i], in[65 + threadIdx.x + i]);
    C2[i] = pack2(in[128 + threadIdx.x + i], in[129 + threadIdx.x + i]);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        if (KIND == 0) A2[i] = fma2(A2[i], C2[i], C2[(i + 1) % NA]);                    // all packed
        if (KIND == 1) A2[i] = fma2(pack2(b[i], b[i]), C2[i], A2[i]);                   // scalar-broadcast multiplicand, acc chain
        if (KIND == 2) A2[i] = fma2(C2[(i + 3) % NA], C2[i], A2[i]);                    // packed multiplicands, acc chain
        if (KIND == 3) A2[i] = add2(A2[i], pack2(b[i], b[i]));                          // scalar-broadcast addend
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NA; ++i) { float x, y; unpack2(A2[i], x, y); s += x + y; }
  if (s == 123.456f) out[0] = s;
}

// experimental variants of pair_gradient_x (klerg_pair.cuh):
//   SC = 1: the accumulations A += (w psi) sc as scalar FFMAs   SC = 2: the dot products too
//   FOLD:   the importance weight folded into the exponent (s2n holds |sc|^2 - log2 w): no multiply, no w operand
template <int D, int WT, int SC, bool FOLD>
__device__ __forceinline__ void pair_gradient_var(const float (&m2x)[WT][D], const float (&x2n)[WT], const u64 (&sc)[D], u64 s2n,
                                                  u64 w2, u64 (&A)[WT][D], u64 (&W)[WT]) {
  float sc0[D], sc1[D];
#pragma unroll
  for (int d = 0; d < D; ++d) unpack2(sc[d], sc0[d], sc1[d]);
  float e0[WT], e1[WT];
  if (SC == 2) {
    float n0, n1;
    unpack2(s2n, n0, n1);
#pragma unroll
    for (int k = 0; k < WT; ++k) { e0[k] = n0 + x2n[k]; e1[k] = n1 + x2n[k]; }
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int k = 0; k < WT; ++k) { e0[k] = fmaf(m2x[k][d], sc0[d], e0[k]); e1[k] = fmaf(m2x[k][d], sc1[d], e1[k]); }
  } else {
    u64 e[WT];
#pragma unroll
    for (int k = 0; k < WT; ++k) e[k] = add2(s2n, pack2(x2n[k], x2n[k]));
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int k = 0; k < WT; ++k) e[k] = fma2(pack2(m2x[k][d], m2x[k][d]), sc[d], e[k]);
#pragma unroll
    for (int k = 0; k < WT; ++k) unpack2(e[k], e0[k], e1[k]);
  }
  float p0[WT], p1[WT];
  float w0, w1;
  unpack2(w2, w0, w1);
#pragma unroll
  for (int k = 0; k < WT; ++k) {
    p0[k] = ex2_neg(e0[k]);
    p1[k] = ex2_neg(e1[k]);
    if (!FOLD) { p0[k] *= w0; p1[k] *= w1; }
  }
#pragma unroll
  for (int k = 0; k < WT; ++k) {
    if (SC >= 1) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float a0, a1;
        unpack2(A[k][d], a0, a1);
        a0 = fmaf(p0[k], sc0[d], a0);
        a1 = fmaf(p1[k], sc1[d], a1);
        A[k][d] = pack2(a0, a1);
      }
    } else {
      const u64 wp = pack2(p0[k], p1[k]);
#pragma unroll
      for (int d = 0; d < D; ++d) A[k][d] = fma2(wp, sc[d], A[k][d]);
    }
    W[k] = add2(W[k], pack2(p0[k], p1[k]));
  }
}

// gradient loop: NW warps, each owns WT states; sweeps `tiles` tiles of TS samples held in shared memory
// FORM 0: difference form (pair_gradient)   1: expanded form (pair_gradient_x, interleaved)
// FORM 2: expanded form, samples NOT reloaded from shared memory (register operands)   3: expanded, no MUFU
template <int D, int WT, int FORM, int NT>
__global__ void __launch_bounds__(NT) grad_loop(int tiles, const float* __restrict__ in, float* out) {
  extern __shared__ __align__(16) float sm[];  // (D + 3) rows of TS floats
  for (int e = threadIdx.x; e < (D + 3) * TS; e += NT) sm[e] = in[e % 3001] * ((e / TS) == D + 2 ? 4.f : 1.f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  u64 xs2[WT][D], acc[WT][D], wacc[WT];
  float m2x[WT][D], x2n[WT];
#pragma unroll
  for (int k = 0; k < WT; ++k) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float x = in[(threadIdx.x >> 5) * 16 + k * 8 + d];
      xs2[k][d] = pack2(x, x);
      m2x[k][d] = -2.f * x;
      acc[k][d] = pack2(0.f, 0.f);
    }
    x2n[k] = in[k] * 3.f;
    wacc[k] = pack2(0.f, 0.f);
  }
  const float* wrow = sm + D * TS;
  const float* nrow = sm + (D + 2) * TS;
  for (int t = 0; t < tiles; ++t) {
    if (FORM == 2 || FORM == 3) {
      u64 s2[D];
#pragma unroll
      for (int d = 0; d < D; ++d) s2[d] = *reinterpret_cast<const u64*>(&sm[d * TS + 2 * lane]);
      u64 w2 = *reinterpret_cast<const u64*>(&wrow[2 * lane]);
      u64 n2 = *reinterpret_cast<const u64*>(&nrow[2 * lane]);
      for (int pb = 0; pb < TS; pb += 64) {
        pair_gradient_x<D, WT, WT>(m2x, x2n, s2, n2, w2, acc, wacc);
        n2 = add2(n2, w2);  // keep the iterations distinct
      }
    } else {
      for (int pb = 0; pb < TS; pb += 64) {
        const int i = pb + 2 * lane;
        u64 s2[D];
#pragma unroll
        for (int d = 0; d < D; ++d) s2[d] = *reinterpret_cast<const u64*>(&sm[d * TS + i]);
        const u64 w2 = *reinterpret_cast<const u64*>(&wrow[i]);
        if (FORM == 0) {
          pair_gradient<D, WT>(xs2, s2, w2, acc, true);
        } else if (FORM >= 4) {
          const u64 n2 = *reinterpret_cast<const u64*>(&nrow[i]);
          // 4: scalar accumulations  5: all scalar  6: packed, w folded  7: scalar accumulations, w folded
          pair_gradient_var<D, WT, FORM == 4 || FORM == 7 ? 1 : (FORM == 5 ? 2 : 0), FORM == 6 || FORM == 7>(
              m2x, x2n, s2, n2, (FORM == 6 || FORM == 7) ? n2 : w2, acc, wacc);
        } else {
          const u64 n2 = *reinterpret_cast<const u64*>(&nrow[i]);
          pair_gradient_x<D, WT, WT>(m2x, x2n, s2, n2, w2, acc, wacc);
        }
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int k = 0; k < WT; ++k) {
    float a, b;
#pragma unroll
    for (int d = 0; d < D; ++d) { unpack2(acc[k][d], a, b); r += a + b; }
    unpack2(wacc[k], a, b);
    r += a + b;
  }
  if (r == 123.456f) out[0] = r;
}

// forward loop in the fused kernel's configuration: samples in registers (P pairs per thread), T state rows in smem
template <int D, int P, int FORM, int NT>
__global__ void __launch_bounds__(NT) fwd_loop(int reps, int T, const float* __restrict__ in, float* out) {
  __shared__ __align__(16) u64 shd[64 * Row2<D>::DP];
  __shared__ __align__(16) float shx[64 * RowX<D>::NF];
  for (int e = threadIdx.x; e < 64 * Row2<D>::DP; e += NT) shd[e] = pack2(in[e % 777], in[e % 777]);
  for (int e = threadIdx.x; e < 64 * RowX<D>::NF; e += NT) shx[e] = in[e % 777];
  __syncthreads();
  u64 s2[D][P], s2n[P], acc[P];
  float emin[2 * P];
#pragma unroll
  for (int q = 0; q < P; ++q) {
#pragma unroll
    for (int d = 0; d < D; ++d) s2[d][q] = pack2(in[threadIdx.x + d * 8 + q], in[threadIdx.x + d * 8 + q + 4]);
    s2n[q] = pack2(in[threadIdx.x + q] * 5.f, in[threadIdx.x + q + 1] * 5.f);
    acc[q] = pack2(0.f, 0.f);
  }
  for (int r = 0; r < reps; ++r) {
    if (FORM == 0)
      pair_forward<D, P, 0>(shd, 0, T, s2, acc, emin);
    else
      pair_forward_x<D, P, 0>(shx, 0, T, s2, s2n, acc, emin);
#pragma unroll
    for (int q = 0; q < P; ++q) s2n[q] = add2(s2n[q], acc[q]);
  }
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < P; ++q) { float a, b; unpack2(acc[q], a, b); s += a + b; }
  if (s == 123.456f) out[0] = s;
}

template <typename F>
static double time_ms(F launch) {
  launch();
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float *in, *out;
  CK(cudaMalloc(&in, 8192 * 4));
  CK(cudaMalloc(&out, 64));
  float h[8192];
  for (int i = 0; i < 8192; ++i) h[i] = 0.25f + 0.001f * (i % 97);
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  const double clk = 1.965e9;
  printf("SMs %d; lane-ops per clock per SM at %.0f MHz\n", sms, clk / 1e6);
  {
    const int iters = 4000, blocks = sms * 4;
    const char* names[4] = {"FFMA2 packed operands", "FFMA2 scalar-broadcast multiplicand", "FFMA2 packed, accumulate chain", "FADD2 scalar-broadcast addend"};
#define PR(K) { double ms = time_ms([&] { probe<K><<<blocks, 256>>>(iters, in, out); }); \
    printf("%-40s %8.3f ms  %7.2f lane-ops/clk/SM\n", names[K], ms, 2.0 * iters * 64 * 256 * blocks / (ms * 1e-3) / clk / sms); }
    PR(0) PR(1) PR(2) PR(3)
  }
  const int tiles = 40;
#define GL(D, WT, FORM, NT, NAME) { \
    auto k = grad_loop<D, WT, FORM, NT>; \
    const size_t smem = sizeof(float) * (D + 3) * TS; \
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    double ms = time_ms([&] { k<<<sms, NT, smem>>>(tiles, in, out); }); \
    const double pairs = (double)tiles * TS * (NT / 32) * WT * sms; \
    const int ops = (FORM == 0) ? (3 * D + 1) : (2 * D + 3); \
    printf("%-52s %8.3f ms  %9.3e pairs/s  FP32 %5.1f%%  MUFU %5.1f%%\n", NAME, ms, pairs / (ms * 1e-3), \
           100.0 * pairs * ops / (ms * 1e-3) / (clk * sms * 128), 100.0 * pairs / (ms * 1e-3) / (clk * sms * 16)); }
  GL(6, 4, 0, 512, "grad D=6 diff     16 warps x 4 states")
  GL(6, 3, 0, 512, "grad D=6 diff     16 warps x 3 states")
  GL(6, 4, 1, 512, "grad D=6 expanded 16 warps x 4 states")
  GL(6, 3, 1, 512, "grad D=6 expanded 16 warps x 3 states")
  GL(6, 4, 4, 512, "grad D=6 expanded 16w x 4, scalar accumulate FFMAs")
  GL(6, 4, 5, 512, "grad D=6 expanded 16w x 4, all scalar FFMAs")
  GL(6, 4, 6, 512, "grad D=6 expanded 16w x 4, w folded into exponent")
  GL(6, 4, 7, 512, "grad D=6 expanded 16w x 4, scalar accumulate + fold")
  GL(6, 3, 4, 512, "grad D=6 expanded 16w x 3, scalar accumulate FFMAs")
  GL(6, 3, 7, 512, "grad D=6 expanded 16w x 3, scalar accumulate + fold")
  GL(6, 4, 1, 384, "grad D=6 expanded 12 warps x 4 states")
  GL(6, 5, 1, 384, "grad D=6 expanded 12 warps x 5 states")
  GL(6, 4, 2, 384, "grad D=6 expanded 12 warps x 4, no smem loads")
  GL(6, 4, 1, 256, "grad D=6 expanded  8 warps x 4 states")
  GL(6, 6, 1, 256, "grad D=6 expanded  8 warps x 6 states")
  GL(3, 5, 0, 640, "grad D=3 diff     20 warps x 5 states")
  GL(3, 5, 1, 640, "grad D=3 expanded 20 warps x 5 states")
  GL(3, 5, 1, 320, "grad D=3 expanded 10 warps x 5 states")
  GL(2, 5, 0, 640, "grad D=2 diff     20 warps x 5 states")
  GL(2, 5, 1, 640, "grad D=2 expanded 20 warps x 5 states")
  const int reps = 2000, T = 50;
#define FL(D, P, FORM, NT, NAME) { \
    double ms = time_ms([&] { fwd_loop<D, P, FORM, NT><<<sms, NT>>>(reps, T, in, out); }); \
    const double pairs = (double)reps * T * NT * 2 * P * sms; \
    const int ops = (FORM == 0) ? (2 * D + 1) : (D + 2); \
    printf("%-52s %8.3f ms  %9.3e pairs/s  FP32 %5.1f%%  MUFU %5.1f%%\n", NAME, ms, pairs / (ms * 1e-3), \
           100.0 * pairs * ops / (ms * 1e-3) / (clk * sms * 128), 100.0 * pairs / (ms * 1e-3) / (clk * sms * 16)); }
  FL(6, 2, 0, 512, "fwd  D=6 diff     512 thr x 4 samples")
  FL(6, 2, 1, 512, "fwd  D=6 expanded 512 thr x 4 samples")
  FL(6, 2, 1, 384, "fwd  D=6 expanded 384 thr x 4 samples")
  FL(6, 1, 1, 512, "fwd  D=6 expanded 512 thr x 2 samples")
  FL(3, 2, 0, 640, "fwd  D=3 diff     640 thr x 4 samples")
  FL(3, 2, 1, 640, "fwd  D=3 expanded 640 thr x 4 samples")
  FL(2, 2, 0, 640, "fwd  D=2 diff     640 thr x 4 samples")
  FL(2, 2, 1, 640, "fwd  D=2 expanded 640 thr x 4 samples")
  return 0;
}

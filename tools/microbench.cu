// Instruction-rate probes for the roofline denominators of the pairwise kernels (sm_100a).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
// Prints lane-ops per clock per SM for scalar/packed FP32 forms, MUFU.EX2 and the pair mixes.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

constexpr int NA = 8;

template <int KIND>
__global__ void __launch_bounds__(256) probe(int iters, const float* __restrict__ in, float* out, long long* cyc) {
  float a[NA], b[NA], c[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[64 + threadIdx.x + i]; c[i] = in[128 + threadIdx.x + i]; }
  uint64_t A2[NA], B2[NA], C2[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) { A2[i] = pk(a[i], a[i] + 1.f); B2[i] = pk(b[i], b[i]); C2[i] = pk(c[i], c[i]); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));
        if (KIND == 1) asm volatile("fma.rn.f32 %0, %0, %1, 0f3A83126F;" : "+f"(a[i]) : "f"(b[i]));
        if (KIND == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        if (KIND == 3) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        if (KIND == 4) A2[i] = fma2(A2[i], B2[i], C2[i]);
        if (KIND == 5) A2[i] = add2(A2[i], B2[i]);
        if (KIND == 6) A2[i] = mul2(A2[i], B2[i]);
        if (KIND == 7) a[i] = ex2(a[i]);
        if (KIND == 8) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(c[(i + 1) % NA]));  // acc += b*c
        if (KIND == 9) asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(a[i]) : "f"(b[i]), "f"(c[(i + r) % NA]));      // independent
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NA; ++i) { float x, y; upk(A2[i], x, y); s += a[i] + x + y; }
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- pair mixes: D=3 forward / gradient inner loops, states broadcast from smem ----
constexpr int TS = 64;  // states per pass

template <int D, int SPT, bool GRAD>
__global__ void __launch_bounds__(256) mix_scalar(int iters, const float* __restrict__ in, float* out, long long* cyc) {
  __shared__ __align__(16) float sh[TS * 8];
  for (int e = threadIdx.x; e < TS * 8; e += 256) sh[e] = in[e % 512] * 3.f;
  __syncthreads();
  float s[D][SPT], acc[SPT], w[SPT], g[D];
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int q = 0; q < SPT; ++q) s[d][q] = in[threadIdx.x + d * 8 + q];
#pragma unroll
  for (int q = 0; q < SPT; ++q) { acc[q] = 0.f; w[q] = in[300 + threadIdx.x + q]; }
#pragma unroll
  for (int d = 0; d < D; ++d) g[d] = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int j = 0; j < TS; ++j) {
      float x[D];
      const float4 v = *reinterpret_cast<const float4*>(&sh[j * 8]);
      x[0] = v.x; if (D > 1) x[1] = v.y; if (D > 2) x[2] = v.z; if (D > 3) x[3] = v.w;
      if (D > 4) { const float4 u = *reinterpret_cast<const float4*>(&sh[j * 8 + 4]); x[4] = u.x; if (D > 5) x[5] = u.y; }
#pragma unroll
      for (int q = 0; q < SPT; ++q) {
        float df[D];
        df[0] = x[0] - s[0][q];
        float e = -df[0] * df[0];
#pragma unroll
        for (int d = 1; d < D; ++d) { df[d] = x[d] - s[d][q]; e = fmaf(-df[d], df[d], e); }
        if (!GRAD) acc[q] += ex2(e);
        else {
          const float wp = w[q] * ex2(e);
#pragma unroll
          for (int d = 0; d < D; ++d) g[d] = fmaf(wp, df[d], g[d]);
        }
      }
    }
  }
  long long t1 = clock64();
  float r = 0.f;
#pragma unroll
  for (int q = 0; q < SPT; ++q) r += acc[q];
#pragma unroll
  for (int d = 0; d < D; ++d) r += g[d];
  if (r == 123.456f) out[0] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// packed: SPT samples per thread as SPT/2 f32x2 pairs; states stored duplicated {x,x} in smem
template <int D, int SPT, bool GRAD>
__global__ void __launch_bounds__(256) mix_packed(int iters, const float* __restrict__ in, float* out, long long* cyc) {
  constexpr int P = SPT / 2;
  __shared__ __align__(16) float sh[TS * 16];
  for (int e = threadIdx.x; e < TS * 16; e += 256) sh[e] = in[(e / 2) % 512] * 3.f;
  __syncthreads();
  uint64_t s[D][P], acc[P], w[P], g[D];
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int q = 0; q < P; ++q) s[d][q] = pk(-in[threadIdx.x + d * 8 + q], -in[threadIdx.x + d * 8 + q + 4]);  // negated samples
#pragma unroll
  for (int q = 0; q < P; ++q) { acc[q] = pk(0.f, 0.f); w[q] = pk(in[300 + threadIdx.x + q], in[304 + threadIdx.x + q]); }
#pragma unroll
  for (int d = 0; d < D; ++d) g[d] = pk(0.f, 0.f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int j = 0; j < TS; ++j) {
      uint64_t x[D];
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&sh[j * 16]);
      x[0] = v.x; if (D > 1) x[1] = v.y;
      if (D > 2) { const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(&sh[j * 16 + 4]); x[2] = u.x; if (D > 3) x[3] = u.y; }
      if (D > 4) { const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(&sh[j * 16 + 8]); x[4] = u.x; if (D > 5) x[5] = u.y; }
#pragma unroll
      for (int q = 0; q < P; ++q) {
        uint64_t df[D];
        df[0] = add2(x[0], s[0][q]);
        uint64_t e = mul2(df[0], df[0]);
#pragma unroll
        for (int d = 1; d < D; ++d) { df[d] = add2(x[d], s[d][q]); e = fma2(df[d], df[d], e); }
        float e0, e1;
        upk(e, e0, e1);
        const uint64_t ps = pk(ex2(-e0), ex2(-e1));
        if (!GRAD) acc[q] = add2(acc[q], ps);
        else {
          const uint64_t wp = mul2(w[q], ps);
#pragma unroll
          for (int d = 0; d < D; ++d) g[d] = fma2(wp, df[d], g[d]);
        }
      }
    }
  }
  long long t1 = clock64();
  float r = 0.f;
#pragma unroll
  for (int q = 0; q < P; ++q) { float a, b; upk(acc[q], a, b); r += a + b; }
#pragma unroll
  for (int d = 0; d < D; ++d) { float a, b; upk(g[d], a, b); r += a + b; }
  if (r == 123.456f) out[0] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, double lane_ops_per_thread, int blocks, int ctas_per_sm_hint) {
  float *in, *out; long long* cyc;
  CK(cudaMalloc(&in, 4096 * 4)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, blocks * 8));
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = 0.25f + 0.001f * (i % 97);
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  launch(in, out, cyc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    launch(in, out, cyc);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  long long* hc = (long long*)malloc(blocks * 8);
  CK(cudaMemcpy(hc, cyc, blocks * 8, cudaMemcpyDeviceToHost));
  double mean = 0; for (int i = 0; i < blocks; ++i) mean += hc[i]; mean /= blocks;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const double total = lane_ops_per_thread * 256.0 * blocks;
  // per-SM per-clock from the in-kernel cycle counter: (ops of resident CTAs) / cycles
  const double per_clk_sm = lane_ops_per_thread * 256.0 * ctas_per_sm_hint / mean;
  printf("%-34s %9.3f ms  %10.3e ops/s  %7.2f ops/clk/SM (clock64, %d CTA/SM)  eff.clk %.0f MHz\n", name, best, total / (best * 1e-3),
         per_clk_sm, ctas_per_sm_hint, total / (best * 1e-3) / (per_clk_sm * sms) / 1e6);
  free(hc); cudaFree(in); cudaFree(out); cudaFree(cyc);
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d\n", sms);
  const int iters = 4000, cps = 4, blocks = sms * cps;
  const double ops = (double)iters * 8 * NA;
#define P(K, NAME, MULT) run(NAME, [&](const float* in, float* out, long long* cyc) { probe<K><<<blocks, 256>>>(iters, in, out, cyc); }, ops * MULT, blocks, cps)
  P(0, "FFMA  a=a*b+c (3 regs)", 1);
  P(1, "FFMA  a=a*b+imm", 1);
  P(8, "FFMA  acc+=b*c (3 distinct)", 1);
  P(2, "FADD  a=a+b", 1);
  P(9, "FADD  a=b-c (indep)", 1);
  P(3, "FMUL  a=a*b", 1);
  P(4, "FFMA2 (lane-ops = 2/instr)", 2);
  P(5, "FADD2", 2);
  P(6, "FMUL2", 2);
  P(7, "MUFU.EX2", 1);
  const int it2 = 400;
#define M(KER, NAME, SPT) run(NAME, [&](const float* in, float* out, long long* cyc) { KER<<<blocks, 256>>>(it2, in, out, cyc); }, (double)it2 * TS * SPT, blocks, cps)
  M((mix_scalar<3, 4, false>), "fwd  D=3 scalar SPT4  [pairs]", 4);
  M((mix_scalar<3, 8, false>), "fwd  D=3 scalar SPT8  [pairs]", 8);
  M((mix_packed<3, 4, false>), "fwd  D=3 packed SPT4  [pairs]", 4);
  M((mix_packed<3, 8, false>), "fwd  D=3 packed SPT8  [pairs]", 8);
  M((mix_scalar<3, 4, true>), "grad D=3 scalar SPT4  [pairs]", 4);
  M((mix_packed<3, 4, true>), "grad D=3 packed SPT4  [pairs]", 4);
  M((mix_packed<3, 8, true>), "grad D=3 packed SPT8  [pairs]", 8);
  M((mix_scalar<6, 4, false>), "fwd  D=6 scalar SPT4  [pairs]", 4);
  M((mix_packed<6, 4, false>), "fwd  D=6 packed SPT4  [pairs]", 4);
  M((mix_scalar<6, 4, true>), "grad D=6 scalar SPT4  [pairs]", 4);
  M((mix_packed<6, 4, true>), "grad D=6 packed SPT4  [pairs]", 4);
  M((mix_scalar<2, 4, false>), "fwd  D=2 scalar SPT4  [pairs]", 4);
  M((mix_packed<2, 8, false>), "fwd  D=2 packed SPT8  [pairs]", 8);
  return 0;
}

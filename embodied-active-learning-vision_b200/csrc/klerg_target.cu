// Target density p(s) of the VAE sensor model, evaluated on the device.
//
// Replaces VAE.pdf_torch (franka_test/scripts/vae/vae.py:244-275):
//   latent = [z || s - shift]  ->  decode = Linear(zd+sd,H1) ReLU Linear(H1,H2) ReLU Linear(H2,out)
//   p(s)   = amax_l exp( mean_z clamp(decode(latent)[:, l], lo, hi) ),  l < ylogvar_dim
//
// The first layer folds the (sample-independent) z part into a bias: h1_k = relu(c_zk + sum_d W1x[k][d] x_d);
// CUDA cores compute it and write it straight into TENSOR MEMORY as the A operand (tcgen05.st).  The second layer -
// the only dense contraction of the whole KL-ergodic path, [N,H1] x [H1,H2] - runs on the tensor cores:
// tcgen05.mma kind::tf32 in its TS form (A from TMEM, B = W2 from shared memory), M = 128 samples per tile,
// fp32 accumulator in TMEM.  TMEM columns [0,256) hold the accumulator, [256,512) a ring of A stages; the H2 outputs
// are processed in passes of <= 256 columns (H2 = 512: two passes, A re-produced per pass - cheap next to the MMAs).
// fp32 parity with the reference (1e-4 relative is the bar, ~5e-6 measured) comes from the 3xTF32 split: every
// operand is hi + lo (W2: round-to-nearest halves, split once per weight refresh; A: hi = top 10 mantissa bits,
// lo = h - hi, exact) and D += A_lo B_hi + A_hi B_lo + A_hi B_hi.
// The third layer only needs its first ylogvar_dim rows: a dot product in the TMEM -> register epilogue, fused
// with bias + ReLU of layer 2, the clamp, the mean over z, exp and amax.
//
// Roles inside a CTA (192 threads, one CTA per SM - it owns the SM's whole TMEM -, persistent over sample tiles):
//   warps 0-3  thread = sample row = TMEM lane: produce the A stages, later drain the accumulator (epilogue); they
//              run one stage into the next pass before draining, so the MMA lane always has work queued
//   warp 4     one lane streams the pre-split W2 stages global -> shared with TMA bulk copies (cp.async.bulk)
//   warp 5     one lane issues the tcgen05.mma's and commits stage / accumulator barriers
// An operand stage = KS = 4 MMA K-steps (32 of the H1 inputs): 12 MMAs per tcgen05.commit (a commit per K-step costs
// ~80 clk of tensor-pipe time each, measured), one 64 KB bulk copy, 64 TMEM columns of A; the A and W2 stages share
// one 3-deep mbarrier ring.  W2 in shared memory uses the canonical no-swizzle K-major layout (8 rows x 16 B core
// matrices): per K-step [part hi|lo][16-byte K chunk 0|1][row][16 B], SBO (8-row group stride) = 128 B,
// LBO (K chunk stride) = rows * 16 B.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "klerg_b200.h"
#include "klerg_common.cuh"
#include "klerg_pair.cuh"
#include "klerg_tc.cuh"

namespace klerg {
namespace {
using namespace tc;

constexpr int TM = 128;            // samples per tile = UMMA M = TMEM lanes
constexpr int THREADS = 192;       // 4 producer/epilogue warps + loader warp + MMA warp
constexpr int KS = 4;              // MMA K-steps (8 tf32 each) per operand stage: 12 MMAs per tcgen05.commit
constexpr int A_COL0 = 256;        // TMEM columns [0,256) accumulator, [256,512) A ring (16*KS columns per stage)
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 256 / (16 * KS);  // ring depth limit: the A ring's TMEM columns
constexpr size_t SMEM_MAX = 227 * 1024;

struct Dims {
  int sd, zd, nz, h1, h2, nl, lp1, lp;
  int halves, ncols;  // the H2 outputs are processed as `halves` accumulator passes of `ncols` <= 256 TMEM columns
};

__host__ __device__ inline size_t align128(size_t x) { return (x + 127) & ~(size_t)127; }
// packed decoder (device): [t1: nz*h1*lp1 f32][e: h2*lp f32][b3: 16 f32][w2s: halves x (h1/8) stages x 64*ncols bytes]
__host__ __device__ inline size_t off_t1(const Dims&) { return 0; }
__host__ __device__ inline size_t off_e(const Dims& d) { return align128(sizeof(float) * (size_t)d.nz * d.h1 * d.lp1); }
__host__ __device__ inline size_t off_b3(const Dims& d) { return off_e(d) + align128(sizeof(float) * (size_t)d.h2 * d.lp); }
__host__ __device__ inline size_t off_w2(const Dims& d) { return off_b3(d) + 128; }
__host__ __device__ inline size_t b_step_bytes(const Dims& d) { return (size_t)64 * d.ncols; }  // one K-step of W2: [hi|lo][chunk][ncols][16 B]
__host__ __device__ inline size_t packed_bytes(const Dims& d) { return off_w2(d) + (size_t)d.halves * (d.h1 / 8) * b_step_bytes(d); }

bool make_dims(int sd, int zd, int nz, int h1, int h2, int nl, Dims& d) {
  if (sd < 1 || sd > 7) { set_error("target decoder: s_dim=%d outside 1..7", sd); return false; }
  if (zd < 0 || nz < 1 || nz > 64) { set_error("target decoder: z_dim=%d / z vectors=%d out of range", zd, nz); return false; }
  if (h1 < 8 || (h1 % 8) || h1 > 1024) { set_error("target decoder: first hidden width %d must be a multiple of 8 in 8..1024", h1); return false; }
  if (h2 < 32 || (h2 % 32) || h2 > 512 || (h2 > 256 && (h2 % 64))) {
    set_error("target decoder: second hidden width %d must be a multiple of 32 in 32..256 or of 64 in 320..512 (TMEM accumulator passes)", h2);
    return false;
  }
  if (nl < 1 || nl > 15) { set_error("target decoder: ylogvar_dim=%d outside 1..15", nl); return false; }
  d.sd = sd; d.zd = zd; d.nz = nz; d.h1 = h1; d.h2 = h2; d.nl = nl;
  d.lp1 = sd <= 3 ? 4 : 8;
  d.lp = nl <= 3 ? 4 : 16;
  d.halves = h2 > 256 ? 2 : 1;
  d.ncols = h2 / d.halves;
  return true;
}


// ---- pack: fold z into the first-layer bias, interleave the epilogue table, split W2 into tf32 hi/lo stages ----
__global__ void pack_decoder_kernel(const Dims d, const float* __restrict__ w1, const float* __restrict__ b1,
                                    const float* __restrict__ z, const float* __restrict__ w2,
                                    const float* __restrict__ b2, const float* __restrict__ w3,
                                    const float* __restrict__ b3, unsigned char* __restrict__ packed) {
  float* t1 = (float*)(packed + off_t1(d));
  float* e = (float*)(packed + off_e(d));
  float* pb3 = (float*)(packed + off_b3(d));
  uint32_t* w2s = (uint32_t*)(packed + off_w2(d));
  const long long n_t1 = (long long)d.nz * d.h1 * d.lp1, n_e = (long long)d.h2 * d.lp, n_w = (long long)d.h1 * d.h2;
  const long long total = n_t1 + n_e + 16 + n_w;
  const int ld1 = d.zd + d.sd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (i < n_t1) {
      const int slot = (int)(i % d.lp1), k = (int)((i / d.lp1) % d.h1), zi = (int)(i / ((long long)d.lp1 * d.h1));
      float v = 0.f;
      if (slot == 0) {
        v = b1[k];
        for (int j = 0; j < d.zd; ++j) v = fmaf(w1[(size_t)k * ld1 + j], z[(size_t)zi * d.zd + j], v);
      } else if (slot <= d.sd) {
        v = w1[(size_t)k * ld1 + d.zd + slot - 1];
      }
      t1[i] = v;
    } else if (i < n_t1 + n_e) {
      const long long r = i - n_t1;
      const int slot = (int)(r % d.lp), j = (int)(r / d.lp);
      float v = 0.f;
      if (slot == 0) v = b2[j];
      else if (slot <= d.nl) v = w3[(size_t)(slot - 1) * d.h2 + j];
      e[r] = v;
    } else if (i < n_t1 + n_e + 16) {
      const int l = (int)(i - n_t1 - n_e);
      pb3[l] = l < d.nl ? b3[l] : 0.f;
    } else {
      // W2[n][k] -> stage (half = n/ncols, ks = k/8): [part][chunk c = (k%8)/4][n % ncols][k%4]
      const long long r = i - n_t1 - n_e - 16;
      const int k = (int)(r % d.h1), n = (int)(r / d.h1);
      const float w = w2[r];
      const uint32_t hi = rna_tf32(w);
      const uint32_t lo = rna_tf32(w - __uint_as_float(hi));
      const int ks = k >> 3, c = (k >> 2) & 1, q = k & 3, hf = n / d.ncols, nn = n % d.ncols;
      const size_t stage_words = (size_t)16 * d.ncols;  // 64*ncols bytes
      const size_t base = ((size_t)hf * (d.h1 >> 3) + ks) * stage_words + ((size_t)c * d.ncols + nn) * 4 + q;
      w2s[base] = hi;
      w2s[base + (size_t)8 * d.ncols] = lo;  // part 1 starts after 2 chunks x ncols rows x 4 words
    }
  }
}

// ---- main kernel ---------------------------------------------------------------------------------------------
struct DecodeArgs {
  const unsigned char* packed;
  const float* samples;
  const float* shift;
  float* out;
  unsigned* fault;
  long long n;
  Dims d;
  float clamp_lo, clamp_hi;
  int stages, tmem_cols;
};

struct SmemLayout {
  size_t b, t1, e, e2, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(const Dims& d, int stages) {
  SmemLayout s;
  size_t o = 0;
  s.b = o;  o += (size_t)stages * KS * b_step_bytes(d);
  s.t1 = o; o = align128(o + sizeof(float) * (size_t)d.nz * d.h1 * d.lp1);
  s.e = o;  o = align128(o + sizeof(float) * (size_t)d.h2 * d.lp);
  s.e2 = o; o = align128(o + (d.nl == 1 ? sizeof(float) * 2 * (size_t)d.h2 : 0));  // single logvar column: {b2 pair, w3 pair}
  s.bars = o; o += 8 * (3 * MAX_STAGES + 2) + 16;  // full_a full_b empty tmem_full tmem_empty | tmem addr, abort
  s.total = o;
  return s;
}


template <int LP1>
__device__ __forceinline__ float first_layer(const float* __restrict__ row, const float (&x)[LP1 - 1]) {
  // row = {c_zk, W1x[k][0..sd-1], 0...}; zero weights make the unused lanes vanish
  const float4 t = *reinterpret_cast<const float4*>(row);
  float v = fmaf(t.y, x[0], t.x);
  v = fmaf(t.z, x[1], v);
  v = fmaf(t.w, x[2], v);
  if constexpr (LP1 == 8) {
    const float4 s = *reinterpret_cast<const float4*>(row + 4);
    v = fmaf(s.x, x[3], v);
    v = fmaf(s.y, x[4], v);
    v = fmaf(s.z, x[5], v);
    v = fmaf(s.w, x[6], v);
  }
  return fmaxf(v, 0.f);
}

template <int LP1, int LP>
__global__ void __launch_bounds__(THREADS, 1) target_decoder_kernel(const DecodeArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Dims d = a.d;
  const SmemLayout sl = smem_layout(d, a.stages);
  unsigned char* sB = smem + sl.b;
  float* s_t1 = (float*)(smem + sl.t1);
  float* s_e = (float*)(smem + sl.e);
  float* s_e2 = (float*)(smem + sl.e2);
  unsigned long long* full_a = (unsigned long long*)(smem + sl.bars);
  unsigned long long* full_b = full_a + MAX_STAGES;
  unsigned long long* empty = full_b + MAX_STAGES;
  unsigned long long* tmem_full = empty + MAX_STAGES;
  unsigned long long* tmem_empty = tmem_full + 1;
  uint32_t* s_tmem = (uint32_t*)(tmem_empty + 1);
  volatile unsigned* s_abort = (volatile unsigned*)(s_tmem + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ksteps = d.h1 >> 3;                 // MMA K-steps (8 tf32 each) per accumulator pass
  const int nst = (ksteps + KS - 1) / KS;       // operand stages per pass (KS K-steps each, the last one may be short)
  const int stages = a.stages;                  // ring depth: W2 stages in shared memory, A stages in TMEM
  const long long tiles = (a.n + TM - 1) / TM;
  const unsigned bstep = (unsigned)b_step_bytes(d);
  const int passes = d.nz * d.halves;           // accumulator passes per sample tile
  const long long my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long n_items = my_tiles * passes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      bar_init(&full_a[s], TM);
      bar_init(&full_b[s], 1);
      bar_init(&empty[s], 1);
    }
    bar_init(tmem_full, 1);
    bar_init(tmem_empty, TM);
    *s_abort = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float4* g1 = (const float4*)(a.packed + off_t1(d));
    const int n1 = d.nz * d.h1 * d.lp1 / 4;
    for (int i = threadIdx.x; i < n1; i += THREADS) ((float4*)s_t1)[i] = __ldg(g1 + i);
    const float4* g2 = (const float4*)(a.packed + off_e(d));
    const int n2 = d.h2 * d.lp / 4;
    for (int i = threadIdx.x; i < n2; i += THREADS) ((float4*)s_e)[i] = __ldg(g2 + i);
    if (d.nl == 1) {  // columns (j, j+1) -> {b2_j, b2_j+1, w3_j, w3_j+1}: one LDS.128 feeds two packed f32x2 operations
      const float* ge = (const float*)(a.packed + off_e(d));
      for (int j = threadIdx.x; j < d.h2; j += THREADS) {
        s_e2[(j >> 1) * 4 + (j & 1)] = __ldg(ge + (size_t)j * d.lp);
        s_e2[(j >> 1) * 4 + 2 + (j & 1)] = __ldg(ge + (size_t)j * d.lp + 1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;  // accumulator: columns [0, ncols); A ring: columns [A_COL0, A_COL0 + stages*16*KS)

  if (warp < 4) {
    // ================= producers of A (TMEM), then epilogue of the same rows =================
    const int row = threadIdx.x;
    const float* b3 = (const float*)(a.packed + off_b3(d));
    const float inv_nz = 1.0f / (float)d.nz;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const long long total_stages = n_items * nst;
    long long produced = 0;
    // producer cursor (kept incrementally: no divisions in the loop): stage within the pass, pass within the tile, tile
    int p_st = 0, p_zh = 0, p_s = 0;
    unsigned p_ph = 0;
    long long p_tile = 0, x_tile = -1;
    int pending = -1;  // A stage whose TMEM stores are issued but not yet signalled
    int e_zh = 0;      // epilogue cursor
    long long e_tile = 0;
    float x[LP1 - 1];
#pragma unroll
    for (int j = 0; j < LP1 - 1; ++j) x[j] = 0.f;
    float y[LP - 1], ysum[LP - 1];
#pragma unroll
    for (int l = 0; l < LP - 1; ++l) y[l] = ysum[l] = 0.f;
    uint32_t acc_phase = 0;

    for (long long item = 0; item < n_items; ++item) {
      // ---- produce: the rest of this pass plus a head start of one stage into the next pass, so the MMA warp has
      //      work queued while these threads drain the accumulator ----
      long long goal = (item + 1) * nst + (stages > 2 ? 1 : 0);
      if (goal > total_stages) goal = total_stages;
      for (; produced < goal; ++produced) {
        const int z = d.halves == 2 ? (p_zh >> 1) : p_zh;
        if (p_tile != x_tile) {  // first stage of a new sample tile: fetch this row's sample
          x_tile = p_tile;
          const long long rg = (blockIdx.x + p_tile * gridDim.x) * TM + row;
#pragma unroll
          for (int j = 0; j < LP1 - 1; ++j) {
            x[j] = 0.f;
            if (rg < a.n && j < d.sd) x[j] = a.samples[rg * d.sd + j] - (a.shift ? a.shift[j] : 0.f);
          }
        }
        const int ks0 = p_st * KS;
        const int nk = min(KS, ksteps - ks0);
        const float* t1z = s_t1 + ((size_t)z * d.h1 + (size_t)ks0 * 8) * LP1;
        const int sa = p_s;
        bar_wait(&empty[sa], p_ph ^ 1u, s_abort);
        tc_fence_after();
        const uint32_t acol = trow + (uint32_t)(A_COL0 + sa * (16 * KS));
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          if (kk < nk) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              // tf32 split by truncation: hi keeps the top 10 mantissa bits, lo = h - hi is exact in fp32 and the tensor
              // core reads its top 10 bits -> |h - hi - lo_tf32| < 2^-20 |h| (cvt.rna is emulated in SASS: ~5x the work)
              const float h = first_layer<LP1>(t1z + (size_t)(kk * 8 + q) * LP1, x);
              hi[q] = __float_as_uint(h) & 0xFFFFE000u;
              lo[q] = __float_as_uint(h - __uint_as_float(hi[q]));
            }
            if (kk == 0 && pending >= 0) {  // the previous stage's TMEM stores had this K-step's arithmetic to land
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              tc_fence_before();
              bar_arrive(&full_a[pending]);
            }
            tmem_st8(acol + kk * 16, hi);
            tmem_st8(acol + kk * 16 + 8, lo);
          }
        }
        pending = sa;
        if (++p_s == stages) {
          p_s = 0;
          p_ph ^= 1u;
        }
        if (++p_st == nst) {
          p_st = 0;
          if (++p_zh == passes) {
            p_zh = 0;
            ++p_tile;
          }
        }
      }
      if (pending >= 0) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        bar_arrive(&full_a[pending]);
        pending = -1;
      }
      // ---- epilogue of (tile, z, half): bias + ReLU of layer 2, dot with the logvar rows of layer 3 ----
      const long long t_local = e_tile;
      const int zh = e_zh;
      const int hf = d.halves == 2 ? (zh & 1) : 0;
      if (++e_zh == passes) {
        e_zh = 0;
        ++e_tile;
      }
      bar_wait(tmem_full, acc_phase, s_abort);
      acc_phase ^= 1u;
      tc_fence_after();
      if (hf == 0) {
#pragma unroll
        for (int l = 0; l < LP - 1; ++l) y[l] = 0.f;
      }
      if (LP == 4 && d.nl == 1) {
        // single logvar column (the reference's default): packed f32x2 arithmetic, two columns per instruction
        u64 y2a = pack2(0.f, 0.f), y2b = pack2(0.f, 0.f);
        for (int c0 = 0; c0 < d.ncols; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(trow + (uint32_t)c0, r);
          tmem_ld_wait();
          const float4* e2 = reinterpret_cast<const float4*>(s_e2) + ((hf * d.ncols + c0) >> 1);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t0 = e2[j >> 1], t1 = e2[(j >> 1) + 1];
            float h0, h1, h2, h3;
            unpack2(add2(pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), pack2(t0.x, t0.y)), h0, h1);
            unpack2(add2(pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), pack2(t1.x, t1.y)), h2, h3);
            y2a = fma2(pack2(t0.z, t0.w), pack2(fmaxf(h0, 0.f), fmaxf(h1, 0.f)), y2a);
            y2b = fma2(pack2(t1.z, t1.w), pack2(fmaxf(h2, 0.f), fmaxf(h3, 0.f)), y2b);
          }
        }
        float s0, s1, s2, s3;
        unpack2(y2a, s0, s1);
        unpack2(y2b, s2, s3);
        y[0] += (s0 + s1) + (s2 + s3);
      } else
      for (int c0 = 0; c0 < d.ncols; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + (uint32_t)c0, r);
        tmem_ld_wait();
        const float* erow = s_e + (size_t)(hf * d.ncols + c0) * LP;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 ev = *reinterpret_cast<const float4*>(erow + j * LP);
          const float hj = fmaxf(__uint_as_float(r[j]) + ev.x, 0.f);
          y[0] = fmaf(ev.y, hj, y[0]);
          y[1] = fmaf(ev.z, hj, y[1]);
          y[2] = fmaf(ev.w, hj, y[2]);
          if constexpr (LP == 16) {
#pragma unroll
            for (int g = 1; g < 4; ++g) {
              const float4 ew = *reinterpret_cast<const float4*>(erow + j * LP + 4 * g);
              y[4 * g - 1] = fmaf(ew.x, hj, y[4 * g - 1]);
              y[4 * g + 0] = fmaf(ew.y, hj, y[4 * g + 0]);
              y[4 * g + 1] = fmaf(ew.z, hj, y[4 * g + 1]);
              y[4 * g + 2] = fmaf(ew.w, hj, y[4 * g + 2]);
            }
          }
        }
      }
      tc_fence_before();
      bar_arrive(tmem_empty);  // accumulator drained: the MMA warp may start the next (tile, z, half)
      if (hf == d.halves - 1) {
#pragma unroll
        for (int l = 0; l < LP - 1; ++l)
          if (l < d.nl) ysum[l] += fminf(fmaxf(y[l] + b3[l], a.clamp_lo), a.clamp_hi);  // torch.clamp (vae.py:266)
      }
      if (zh == passes - 1) {
        float p = -INFINITY;
#pragma unroll
        for (int l = 0; l < LP - 1; ++l) {
          if (l < d.nl) p = fmaxf(p, expf(d.nz > 1 ? ysum[l] * inv_nz : ysum[l]));  // mean over z, exp, amax (vae.py:267-273)
          ysum[l] = 0.f;
        }
        const long long rg = (blockIdx.x + t_local * gridDim.x) * TM + row;
        if (rg < a.n) a.out[rg] = p;
      }
    }
  } else if (warp == 4) {
    // ================= W2 stages: one TMA bulk copy per stage (KS K-steps, contiguous in the packed buffer) =================
    if (lane == 0) {
      const unsigned char* w2s = a.packed + off_w2(d);
      int s = 0, zh = 0;
      unsigned ph = 0;
      for (long long item = 0; item < n_items; ++item) {
        const int hf = d.halves == 2 ? (zh & 1) : 0;
        if (++zh == passes) zh = 0;
        for (int st = 0; st < nst; ++st) {
          const int ks0 = st * KS;
          const unsigned bytes = (unsigned)min(KS, ksteps - ks0) * bstep;
          bar_wait(&empty[s], ph ^ 1u, s_abort);
          bar_expect_tx(&full_b[s], bytes);
          bulk_g2s(sB + (size_t)s * (KS * bstep), w2s + ((size_t)hf * ksteps + ks0) * bstep, bytes, &full_b[s]);
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else {
    // ================= MMA issue: A from TMEM, B from shared memory; one commit per stage =================
    if (lane == 0) {
      const uint32_t idesc = instr_desc_tf32(TM, d.ncols);
      // B: LBO = stride between the two 16-byte K chunks of an MMA, SBO = stride between 8-row groups
      const uint32_t b_lbo = (uint32_t)d.ncols * 16, b_sbo = 128;
      int s = 0;
      unsigned ph = 0;
      for (long long item = 0; item < n_items; ++item) {
        if (item > 0) bar_wait(tmem_empty, (unsigned)(item - 1) & 1u, s_abort);
        tc_fence_after();
        for (int st = 0; st < nst; ++st) {
          const int nk = min(KS, ksteps - st * KS);
          bar_wait(&full_a[s], ph, s_abort);
          bar_wait(&full_b[s], ph, s_abort);
          tc_fence_after();
          const uint32_t a0 = tmem + (uint32_t)(A_COL0 + s * (16 * KS));
          const uint32_t b0 = smem_addr(sB + (size_t)s * (KS * bstep));
          for (int kk = 0; kk < nk; ++kk) {
            const uint32_t a_hi = a0 + kk * 16, a_lo = a_hi + 8;
            const uint32_t b_hi = b0 + kk * bstep, b_lo = b_hi + 32u * (uint32_t)d.ncols;
            const uint64_t db_hi = smem_desc(b_hi, b_lbo, b_sbo), db_lo = smem_desc(b_lo, b_lbo, b_sbo);
            tc_mma_tf32_ts(tmem, a_lo, db_hi, idesc, (st | kk) ? 1u : 0u);  // small terms first
            tc_mma_tf32_ts(tmem, a_hi, db_lo, idesc, 1u);
            tc_mma_tf32_ts(tmem, a_hi, db_hi, idesc, 1u);
          }
          tc_commit(&empty[s]);  // A and W2 stage free once these MMAs have read their operands
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        tc_commit(tmem_full);  // accumulator complete -> epilogue
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && *s_abort && a.fault) atomicExch(a.fault, 1u);
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

template <int LP1, int LP>
int launch_decoder(const DecodeArgs& a, int grid, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(target_decoder_kernel<LP1, LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
    if (e != cudaSuccess) { set_error("target_decoder: smem opt-in failed: %s", cudaGetErrorString(e)); return -4; }
    configured = true;
  }
  target_decoder_kernel<LP1, LP><<<grid, THREADS, smem, st>>>(a);
  return check_launch("target_decoder_kernel");
}

}  // namespace
}  // namespace klerg

using namespace klerg;

extern "C" size_t klerg_target_decoder_packed_bytes(int32_t s_dim, int32_t z_dim, int32_t n_z, int32_t h1, int32_t h2,
                                                    int32_t n_logvar) {
  Dims d;
  if (!make_dims(s_dim, z_dim, n_z, h1, h2, n_logvar, d)) return 0;
  return packed_bytes(d);
}

extern "C" int klerg_target_decoder_pack(const float* w1, const float* b1, const float* z, const float* w2,
                                         const float* b2, const float* w3, const float* b3, int32_t s_dim,
                                         int32_t z_dim, int32_t n_z, int32_t h1, int32_t h2, int32_t n_logvar,
                                         void* packed, void* stream) {
  Dims d;
  if (!make_dims(s_dim, z_dim, n_z, h1, h2, n_logvar, d)) return -1;
  if (!w1 || !b1 || (!z && z_dim > 0) || !w2 || !b2 || !w3 || !b3 || !packed) { set_error("target_decoder_pack: null pointer"); return -1; }
  if ((uintptr_t)packed & 127) { set_error("target_decoder_pack: packed buffer must be 128-byte aligned"); return -1; }
  pack_decoder_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(d, w1, b1, z, w2, b2, w3, b3, (unsigned char*)packed);
  return check_launch("pack_decoder_kernel");
}

extern "C" int klerg_target_decoder_pdf(const void* packed, int32_t s_dim, int32_t z_dim, int32_t n_z, int32_t h1,
                                        int32_t h2, int32_t n_logvar, const float* samples, int64_t N,
                                        const float* shift, float clamp_lo, float clamp_hi, float* p_out,
                                        uint32_t* fault, void* stream) {
  Dims d;
  if (!make_dims(s_dim, z_dim, n_z, h1, h2, n_logvar, d)) return -1;
  if (N < 0) { set_error("target_decoder_pdf: N < 0"); return -1; }
  if (N == 0) return 0;
  if (!packed || !samples || !p_out) { set_error("target_decoder_pdf: null pointer"); return -1; }
  if ((uintptr_t)packed & 127) { set_error("target_decoder_pdf: packed buffer must be 128-byte aligned"); return -1; }
  int stages = MAX_STAGES;
  const int per_sm = 1;  // the CTA owns the SM's whole TMEM (accumulator + A ring)
  while (stages >= 2 && smem_layout(d, stages).total > SMEM_MAX) --stages;
  if (stages < 2) { set_error("target_decoder_pdf: decoder tables (%d z vectors x %d) do not fit in shared memory", n_z, h1); return -2; }
  DecodeArgs a;
  a.packed = (const unsigned char*)packed;
  a.samples = samples;
  a.shift = shift;
  a.out = p_out;
  a.fault = fault;
  a.n = N;
  a.d = d;
  a.clamp_lo = clamp_lo;
  a.clamp_hi = clamp_hi;
  a.stages = stages;
  a.tmem_cols = TMEM_COLS;
  const long long tiles = (N + TM - 1) / TM;
  const long long slots = (long long)per_sm * sm_count();
  const int grid = (int)(tiles < slots ? tiles : slots);
  const size_t smem = smem_layout(d, stages).total;
  cudaStream_t st = (cudaStream_t)stream;
  if (d.lp1 == 4 && d.lp == 4) return launch_decoder<4, 4>(a, grid, smem, st);
  if (d.lp1 == 8 && d.lp == 4) return launch_decoder<8, 4>(a, grid, smem, st);
  if (d.lp1 == 4 && d.lp == 16) return launch_decoder<4, 16>(a, grid, smem, st);
  return launch_decoder<8, 16>(a, grid, smem, st);
}

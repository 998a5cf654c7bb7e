// Shared-psi KL gradient for K belief targets over one workspace / trajectory (BASELINE config 5, fingerprint test mode).
//
// kldiv_grad_vec (control_torch/klerg_utils.py:12-15,31-36) for all H planner states and K targets p_k:
//   dgdx_k[t][d] = sum_i w_ki * (-(x_td - s_id)/|std_d|) * psi(x_t, s_i),   w_ki = p_ki / q_i  (klerg.py:436)
// psi does not depend on the target, so with K targets the sum over the samples is a contraction
//   S[(k,d')][t] = sum_i V[(k,d')][i] * psi[t][i],  V[(k,d)][i] = w_ki (s'_id - c_d),  V[(k,D)][i] = w_ki
//   dgdx_k[t][d] = gfac_d * ((x'_td - c_d) S[(k,D)][t] - S[(k,d)][t])
// (primes: coordinates pre-scaled so that psi = 2^-|x'-s'|^2; c = centre of the trajectory, which keeps the
// cancellation in the last line small) - and it runs on the tensor cores: tcgen05.mma kind::tf32, M = 128 rows (k,d'),
// N = H (padded to 32 or 64) columns, K-dimension = samples, 8 per instruction, 3xTF32 split for fp32-grade sums,
// one fp32 accumulator in TMEM for the CTA's whole sample range.  The pair kernels recompute psi per target
// (K * (3D+1) FP32 lane-ops per state-sample pair); here psi costs one forward-style evaluation per pair.
//
// Roles inside a CTA (512 threads, one CTA per SM, a contiguous range of samples per CTA, 64 samples per stage - the
// hand-offs between the roles, not their arithmetic, are what a stage costs):
//   warps 8-13 stagers, two per ring slot (the global loads of a stage are issued a ring turn ahead): thread = sample: importance-ratio
//              factor maxc/c_i, centred scaled sample, the K target values -> shared memory; the per-target KL terms
//              (sum p (log p - log c), sum c) ride along
//   warps 0-3  thread = row (k,d') = TMEM lane: V for the stage's samples (hi/lo tf32 halves) -> TMEM (A operand); with
//              <= 64 rows every row sits on two lanes that split the samples
//   warps 4-7  thread = (state t, every other K-step of 8 samples): psi -> shared memory (B operand, no-swizzle K-major)
//   warp 15    one lane issues 24 MMAs per stage and commits
// At the end warps 0-3 drain the accumulator into a per-CTA partial; a small second kernel adds the partials in CTA
// order (doubles) and applies the last line above.
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

constexpr int TG_THREADS = 512;         // warps 0-3 V rows, 4-7 psi, 8-13 stagers (two per ring slot), 15 MMA issue
constexpr int TG_KS = 8;                 // MMA K-steps (8 samples each) per stage: the per-stage hand-offs are the bound
constexpr int TG_SPS = 8 * TG_KS;        // samples per stage
constexpr int TG_ROW = TG_SPS + 4;       // floats per staged row: rows of different targets start 4 banks apart, so the
                                         // V producers' per-target LDS.128 do not pile onto the same banks
constexpr int TG_STAGES = 3;             // ring depth (a deeper ring bought nothing: the producers' issue rate is the limit)
constexpr int TG_ACOL0 = 64;             // TMEM: accumulator columns [0, Hp <= 64), A ring [64, 64 + 3*64)
constexpr int TG_STAGER0 = 8;            // first stager warp; two stager warps per ring slot (32 samples each)
constexpr int TG_NST = 2 * TG_STAGES;    // stager warps
constexpr int TG_MAXK = 32;

struct TGArgs {
  KernelDev k;
  const float* states;  // [H][S] planner states (pre-step trajectory)
  int H, Hp;            // Hp = H padded to 32 or 64 (MMA N, accumulator columns)
  const float* packed;  // [D][ld] pre-scaled samples
  long long N, ld;
  const float* v;       // q_base + q_iter
  const double* totals; // [world][1][2] {sum v, max v} per rank
  int world;
  float floor;
  const float* P;       // [K][p_stride]
  int K, R;             // R = K*(D+1) rows in use
  long long p_stride;
  float* part;          // [grid][128][Hp]
  float* klpart;        // [grid][TG_STAGES][K][2] (one row per stager warp)
  double* grad_out;     // [K][H][D]
  double* kl_out;       // [K][2]
  unsigned* fault;
  long long chunk;      // samples per CTA (multiple of TG_SPS)
};

__host__ __device__ inline size_t tg_slot_floats(int D, int K) { return (size_t)(D + 2 + K) * TG_ROW; }  // D samples rows, r, K targets, ones
__host__ __device__ inline size_t tg_bstage_bytes(int Hp) { return (size_t)TG_KS * 64 * Hp; }
struct TGSmem {
  size_t b, slot, xc, bars, total;
};
__host__ __device__ inline TGSmem tg_smem(int D, int K, int Hp) {
  TGSmem s;
  size_t o = 0;
  s.b = o;    o += TG_STAGES * tg_bstage_bytes(Hp);
  s.slot = o; o += TG_STAGES * sizeof(float) * tg_slot_floats(D, K);
  s.xc = o;   o += sizeof(float) * (64 * KLERG_MAX_D + KLERG_MAX_D);  // centred scaled states [64][D], centre [D]
  o = (o + 127) & ~(size_t)127;
  s.bars = o; o += 8 * (4 * TG_STAGES + 1) + 16;
  s.total = o;
  return s;
}

// KMAX: compile-time bound of the stagers' per-target loops (iterations beyond K would still issue, predicated off)
template <int D, int KMAX>
__global__ void __launch_bounds__(TG_THREADS, 1) targets_gradient_kernel(const TGArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TGSmem sl = tg_smem(D, a.K, a.Hp);
  unsigned char* sB = smem + sl.b;
  float* s_slot = (float*)(smem + sl.slot);
  float* s_xc = (float*)(smem + sl.xc);   // [64][D]
  float* s_c = s_xc + 64 * D;             // [D]
  unsigned long long* full_s = (unsigned long long*)(smem + sl.bars);
  unsigned long long* full_a = full_s + TG_STAGES;
  unsigned long long* full_b = full_a + TG_STAGES;
  unsigned long long* empty = full_b + TG_STAGES;
  unsigned long long* acc_full = empty + TG_STAGES;
  uint32_t* s_tmem = (uint32_t*)(acc_full + 1);
  volatile unsigned* s_abort = (volatile unsigned*)(s_tmem + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = a.H, Hp = a.Hp, K = a.K;
  const size_t slot_f = tg_slot_floats(D, K);
  const unsigned bstage = (unsigned)tg_bstage_bytes(Hp), bstep = 64u * (unsigned)Hp;
  const long long lo = (long long)blockIdx.x * a.chunk;
  long long hi = lo + a.chunk;
  if (hi > a.N) hi = a.N;
  const int n_st = hi > lo ? (int)((hi - lo + TG_SPS - 1) / TG_SPS) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TG_STAGES; ++s) {
      bar_init(&full_s[s], 64);  // the slot's two stager warps
      bar_init(&full_a[s], 128);
      bar_init(&full_b[s], 128);
      bar_init(&empty[s], 1);
    }
    bar_init(acc_full, 1);
    *s_abort = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 15) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // scaled states, then their centre (same arithmetic in every CTA -> bit-identical)
  for (int e = threadIdx.x; e < 64 * D; e += TG_THREADS) {
    const int t = e / D, d = e - t * D;
    s_xc[e] = t < H ? a.states[(size_t)t * a.k.S + a.k.explr[d]] * a.k.a[d] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < D) {
    float mn = s_xc[threadIdx.x], mx = mn;
    for (int t = 1; t < H; ++t) {
      mn = fminf(mn, s_xc[t * D + threadIdx.x]);
      mx = fmaxf(mx, s_xc[t * D + threadIdx.x]);
    }
    s_c[threadIdx.x] = 0.5f * (mn + mx);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * D; e += TG_THREADS) {
    const int t = e / D, d = e - t * D;
    if (t < H) s_xc[e] -= s_c[d];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp >= TG_STAGER0 && warp < TG_STAGER0 + TG_NST) {
    // ================= stagers: warps 8+2q, 9+2q own ring slot q (stages q, q+3, ...), one sample per thread; the
    // global loads of a stage are issued a ring turn ahead =================
    double vsum, vmax;
    gather_totals(a.totals, a.world, 1, 0, vsum, vmax);
    const float vsum_f = (float)vsum;  // divide by the fp32 sum like the reference
    const float maxc_f = (float)fmax(vmax / vsum, (double)a.floor);
    float c_d[D];
#pragma unroll
    for (int d = 0; d < D; ++d) c_d[d] = s_c[d];
    float kl_a[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) kl_a[k] = 0.f;
    float kl_c = 0.f;
    const bool want_kl = a.kl_out != nullptr;  // the per-target KL terms cost a log per target and sample
    const int s = (warp - TG_STAGER0) >> 1, half = (warp - TG_STAGER0) & 1, col = half * 32 + lane;
    unsigned ph = 0;
    // The global loads of a stage are issued one ring turn ahead (into registers), so their latency overlaps the
    // wait for the slot instead of following it.
    float pv[KMAX], sv[D], vi = 1.f;
    bool valid = false;
    auto fetch = [&](int st) {
      const long long i = lo + (long long)st * TG_SPS + col;
      valid = st < n_st && i < hi;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) pv[k] = (k < K && valid) ? __ldg(a.P + (size_t)k * a.p_stride + i) : 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) sv[d] = valid ? __ldg(a.packed + (size_t)d * a.ld + i) : c_d[d];
      vi = valid ? __ldg(a.v + i) : 1.f;
    };
    fetch(s);
    for (int st = s; st < n_st; st += TG_STAGES, ph ^= 1u) {
      bar_wait(&empty[s], ph ^ 1u, s_abort);
      float* slot = s_slot + (size_t)s * slot_f;
      float r = 0.f, logc = 0.f;
      if (valid) {
        const float c = fmaxf(__fdividef(vi, vsum_f), a.floor);
        r = __fdividef(maxc_f, c);  // p * r = p / q with q = c / max c  (klerg.py:436)
        if (want_kl) {
          logc = __logf(c);
          kl_c += c;
        }
      }
#pragma unroll
      for (int d = 0; d < D; ++d) slot[d * TG_ROW + col] = sv[d] - c_d[d];
      slot[D * TG_ROW + col] = r;
      slot[(D + 1 + K) * TG_ROW + col] = 1.f;  // multiplier of the weight rows (d' = D)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          slot[(D + 1 + k) * TG_ROW + col] = pv[k];
          if (want_kl && valid) kl_a[k] += pv[k] * (__logf(pv[k]) - logc);
        }
      }
      bar_arrive(&full_s[s]);
      fetch(st + TG_STAGES);
    }
    kl_c = warp_sum_f(kl_c);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        const float sa = warp_sum_f(kl_a[k]);
        if (lane == 0) {
          a.klpart[(((size_t)blockIdx.x * TG_NST + (warp - TG_STAGER0)) * K + k) * 2 + 0] = sa;
          a.klpart[(((size_t)blockIdx.x * TG_NST + (warp - TG_STAGER0)) * K + k) * 2 + 1] = kl_c;
        }
      }
    }
  } else if (warp < 4) {
    // ================= V producers (A operand in TMEM), later the epilogue =================
    // R <= 64: every row lives on two lanes (l and l + 64), each computing half of a stage's samples (zeros for the
    // other half) - all four warps share the work; the reduce kernel adds the two accumulator rows
    const int reps = a.R <= 64 ? 2 : 1, rows_per_rep = 128 / reps;
    const int row = threadIdx.x % rows_per_rep, myrep = threadIdx.x / rows_per_rep;
    const bool active = row < a.R;
    const int k = active ? row / (D + 1) : 0, dd = active ? row - k * (D + 1) : D;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    int s = 0;
    unsigned ph = 0;
    for (int st = 0; st < n_st; ++st) {
      bar_wait(&full_s[s], ph, s_abort);
      tc_fence_after();
      const float* slot = s_slot + (size_t)s * slot_f;
      // rows beyond R repeat row 0's arithmetic (finite values; their accumulator rows are never read)
      const ulonglong2* rr = reinterpret_cast<const ulonglong2*>(slot + D * TG_ROW);
      const ulonglong2* pp = reinterpret_cast<const ulonglong2*>(slot + (D + 1 + k) * TG_ROW);
      const ulonglong2* ss = reinterpret_cast<const ulonglong2*>(slot + (dd < D ? dd : D + 1 + K) * TG_ROW);
      const uint32_t acol = trow + (uint32_t)(TG_ACOL0 + s * (16 * TG_KS));
#pragma unroll
      for (int kk = 0; kk < TG_KS; ++kk) {
        uint32_t hiw[8], low[8];
        if (reps == 2 && (kk / (TG_KS / 2)) != myrep) {  // the other lane of this row covers these samples
#pragma unroll
          for (int q = 0; q < 8; ++q) hiw[q] = low[q] = 0u;
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const ulonglong2 r4 = rr[kk * 2 + h], p4 = pp[kk * 2 + h], s4 = ss[kk * 2 + h];
            // packed f32x2: val = (p * r) * s', tf32 split by truncation (hi = top 10 mantissa bits, lo = val - hi exact)
            const u64 v0 = mul2(mul2(p4.x, r4.x), s4.x), v1 = mul2(mul2(p4.y, r4.y), s4.y);
            const u64 h0 = v0 & 0xFFFFE000FFFFE000ull, h1 = v1 & 0xFFFFE000FFFFE000ull;
            const u64 l0 = sub2(v0, h0), l1 = sub2(v1, h1);
            hiw[h * 4 + 0] = (uint32_t)h0; hiw[h * 4 + 1] = (uint32_t)(h0 >> 32);
            hiw[h * 4 + 2] = (uint32_t)h1; hiw[h * 4 + 3] = (uint32_t)(h1 >> 32);
            low[h * 4 + 0] = (uint32_t)l0; low[h * 4 + 1] = (uint32_t)(l0 >> 32);
            low[h * 4 + 2] = (uint32_t)l1; low[h * 4 + 3] = (uint32_t)(l1 >> 32);
          }
        }
        tmem_st8(acol + kk * 16, hiw);
        tmem_st8(acol + kk * 16 + 8, low);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      bar_arrive(&full_a[s]);
      if (++s == TG_STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
    // ---- epilogue: this row of the accumulator -> the CTA's partial ----
    float* out = a.part + ((size_t)blockIdx.x * 128 + threadIdx.x) * Hp;  // accumulator row = TMEM lane
    if (n_st > 0) {
      bar_wait(acc_full, 0u, s_abort);
      tc_fence_after();
      for (int c0 = 0; c0 < Hp; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + (uint32_t)c0, r);
        tmem_ld_wait();
        if (active) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(out + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                    __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
    } else if (active) {
      for (int j = 0; j < Hp; ++j) out[j] = 0.f;
    }
  } else if (warp < TG_STAGER0) {
    // ================= psi producers (B operand in shared memory) =================
    const int tid2 = threadIdx.x - 128;
    const int t = tid2 & 63, grp = tid2 >> 6;  // state; the two groups take alternate K-steps (8 samples each)
    const bool row_ok = t < Hp, live = t < H;
    u64 x2[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x2[d] = pack2(s_xc[t * D + d], s_xc[t * D + d]);
    const float live_f = live ? 1.f : 0.f;
    int s = 0;
    unsigned ph = 0;
    for (int st = 0; st < n_st; ++st) {
      bar_wait(&full_s[s], ph, s_abort);
      const float* slot = s_slot + (size_t)s * slot_f;
      unsigned char* bst = sB + (size_t)s * bstage;
#pragma unroll
      for (int kk = grp; kk < TG_KS; kk += 2) {
        uint32_t hiw[8], low[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          u64 e01 = pack2(0.f, 0.f), e23 = e01;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const ulonglong2 s4 = *reinterpret_cast<const ulonglong2*>(slot + d * TG_ROW + kk * 8 + h * 4);
            const u64 d01 = sub2(x2[d], s4.x), d23 = sub2(x2[d], s4.y);
            e01 = fma2(d01, d01, e01);
            e23 = fma2(d23, d23, e23);
          }
          float e[4];
          unpack2(e01, e[0], e[1]);
          unpack2(e23, e[2], e[3]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float psi = ex2_neg(e[q]) * live_f;
            hiw[h * 4 + q] = __float_as_uint(psi) & 0xFFFFE000u;
            low[h * 4 + q] = __float_as_uint(psi - __uint_as_float(hiw[h * 4 + q]));
          }
        }
        if (row_ok) {
          // K-step block: [hi|lo][16-byte sample chunk 0|1][t][16 B]
          unsigned char* base = bst + (size_t)kk * bstep + (size_t)t * 16;
          *reinterpret_cast<uint4*>(base) = make_uint4(hiw[0], hiw[1], hiw[2], hiw[3]);
          *reinterpret_cast<uint4*>(base + 16 * Hp) = make_uint4(hiw[4], hiw[5], hiw[6], hiw[7]);
          *reinterpret_cast<uint4*>(base + 32 * Hp) = make_uint4(low[0], low[1], low[2], low[3]);
          *reinterpret_cast<uint4*>(base + 48 * Hp) = make_uint4(low[4], low[5], low[6], low[7]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bar_arrive(&full_b[s]);
      if (++s == TG_STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
  } else if (warp == 15) {
    // ================= MMA issue =================
    if (lane == 0) {
      const uint32_t idesc = instr_desc_tf32(128, Hp);
      const uint32_t b_lbo = (uint32_t)Hp * 16, b_sbo = 128;
      int s = 0;
      unsigned ph = 0;
      for (int st = 0; st < n_st; ++st) {
        bar_wait(&full_a[s], ph, s_abort);
        bar_wait(&full_b[s], ph, s_abort);
        tc_fence_after();
        const uint32_t a0 = tmem + (uint32_t)(TG_ACOL0 + s * (16 * TG_KS));
        const uint32_t b0 = smem_addr(sB + (size_t)s * bstage);
#pragma unroll
        for (int kk = 0; kk < TG_KS; ++kk) {
          const uint32_t a_hi = a0 + kk * 16, a_lo = a_hi + 8;
          const uint32_t b_hi = b0 + kk * bstep, b_lo = b_hi + 32u * (uint32_t)Hp;
          const uint64_t db_hi = smem_desc(b_hi, b_lbo, b_sbo), db_lo = smem_desc(b_lo, b_lbo, b_sbo);
#ifdef TG_DEBUG_ONE_MMA
          tc_mma_tf32_ts(tmem, a_hi, db_hi, idesc, (st | kk) ? 1u : 0u);
#else
          tc_mma_tf32_ts(tmem, a_lo, db_hi, idesc, (st | kk) ? 1u : 0u);
          tc_mma_tf32_ts(tmem, a_hi, db_lo, idesc, 1u);
          tc_mma_tf32_ts(tmem, a_hi, db_hi, idesc, 1u);
#endif
        }
        tc_commit(&empty[s]);
        if (++s == TG_STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (n_st > 0) tc_commit(acc_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && *s_abort && a.fault) atomicExch(a.fault, 1u);
  if (warp == 15) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// Partials -> grad_out[K][H][D] (doubles, the layout klerg_adjoint takes per target) and kl_out[K][2].
// One CTA per target; thread = (row d' of the target, state t, part of the CTA list): every load is a coalesced
// run of accumulator columns, partial sums are combined in a fixed order (deterministic).
template <int D>
__host__ __device__ constexpr int tg_reduce_parts() { return D + 1 <= 4 ? 4 : 2; }  // CTA = (D+1) rows x 64 states x parts <= 1024 threads

template <int D>
__global__ void __launch_bounds__((D + 1) * 64 * tg_reduce_parts<D>()) targets_reduce_kernel(const TGArgs a, int nblk) {
  constexpr int PARTS = tg_reduce_parts<D>();
  __shared__ float s_x[64 * D], s_c[D];
  __shared__ double s_sum[PARTS][D + 1][64];
  const int k = blockIdx.x;
  for (int e = threadIdx.x; e < 64 * D; e += blockDim.x) {
    const int t = e / D, d = e - t * D;
    s_x[e] = t < a.H ? a.states[(size_t)t * a.k.S + a.k.explr[d]] * a.k.a[d] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < D) {
    float mn = s_x[threadIdx.x], mx = mn;
    for (int t = 1; t < a.H; ++t) {
      mn = fminf(mn, s_x[t * D + threadIdx.x]);
      mx = fmaxf(mx, s_x[t * D + threadIdx.x]);
    }
    s_c[threadIdx.x] = 0.5f * (mn + mx);
  }
  const int t = threadIdx.x & 63, dd = (threadIdx.x >> 6) % (D + 1), part4 = threadIdx.x / (64 * (D + 1));
  const int reps = a.R <= 64 ? 2 : 1;
  double acc = 0.0;
  if (t < a.Hp) {
    const int per = (nblk + PARTS - 1) / PARTS, b0 = part4 * per, b1 = min(nblk, b0 + per);
    for (int b = b0; b < b1; ++b)
      for (int rp = 0; rp < reps; ++rp)
        acc += (double)__ldcg(a.part + ((size_t)b * 128 + (size_t)rp * 64 + (size_t)k * (D + 1) + dd) * a.Hp + t);
  }
  s_sum[part4][dd][t] = acc;
  __syncthreads();
  if (part4 == 0 && dd < D && t < a.H) {
    double sd = 0.0, s0 = 0.0;
#pragma unroll
    for (int q = 0; q < PARTS; ++q) {
      sd += s_sum[q][dd][t];
      s0 += s_sum[q][D][t];
    }
    const double xc = (double)(s_x[t * D + dd] - s_c[dd]);
    a.grad_out[((size_t)k * a.H + t) * D + dd] = (double)a.k.gfac[dd] * (xc * s0 - sd);
  }
  if (a.kl_out && threadIdx.x < 64) {  // warps 0-1: the two KL terms of this target
    const int e = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double kacc = 0.0;
    for (int b = lane; b < nblk * TG_NST; b += 32) kacc += (double)__ldcg(a.klpart + ((size_t)b * a.K + k) * 2 + e);
    kacc = warp_reduce(RED_SUM, kacc);
    if (lane == 0) a.kl_out[k * 2 + e] = kacc;
  }
}

template <int D>
int launch_targets(TGArgs& a, cudaStream_t st) {
  static bool configured = false;
  const TGSmem sl = tg_smem(D, a.K, a.Hp);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(targets_gradient_kernel<D, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(targets_gradient_kernel<D, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(targets_gradient_kernel<D, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("kl_gradient_targets: smem opt-in failed: %s", cudaGetErrorString(e)); return -4; }
    configured = true;
  }
  const int grid = sm_count();
  long long chunk = (a.N + grid - 1) / grid;
  chunk = (chunk + TG_SPS - 1) / TG_SPS * TG_SPS;
  if (chunk < TG_SPS) chunk = TG_SPS;
  a.chunk = chunk;
  if (a.K <= 8) targets_gradient_kernel<D, 8><<<grid, TG_THREADS, sl.total, st>>>(a);
  else if (a.K <= 16) targets_gradient_kernel<D, 16><<<grid, TG_THREADS, sl.total, st>>>(a);
  else targets_gradient_kernel<D, 32><<<grid, TG_THREADS, sl.total, st>>>(a);
  int rc = check_launch("targets_gradient_kernel");
  if (rc) return rc;
  targets_reduce_kernel<D><<<a.K, (D + 1) * 64 * tg_reduce_parts<D>(), 0, st>>>(a, grid);
  return check_launch("targets_reduce_kernel");
}

}  // namespace
}  // namespace klerg

using namespace klerg;

extern "C" size_t klerg_kl_gradient_targets_scratch_bytes(int64_t H, int64_t K) {
  const size_t Hp = H <= 32 ? 32 : 64;
  return (size_t)sm_count() * (128 * Hp + (size_t)TG_NST * K * 2) * sizeof(float);
}

extern "C" int klerg_kl_gradient_targets(const klerg_kernel_spec* k, const float* states, int64_t H,
                                         const float* packed, int64_t N, int64_t ld, const float* v,
                                         const double* totals, int world, const float* P, int64_t K,
                                         int64_t p_stride, float floor, double* grad_part, double* kl_part,
                                         void* scratch, uint32_t* fault, void* stream) {
  KernelDev kd;
  if (!make_kernel_dev(k, kd)) return -1;
  if (H < 1 || H > 64) { set_error("kl_gradient_targets: H must be in 1..64 (accumulator columns of this kernel)"); return -2; }
  if (K < 1 || K > TG_MAXK || K * (kd.D + 1) > 128) { set_error("kl_gradient_targets: K*(D+1) must be <= 128 rows (K <= 32)"); return -2; }
  if (N < 0 || ld < N || (ld & 3) || p_stride < N) { set_error("kl_gradient_targets: bad sizes"); return -1; }
  if (!states || !packed || !v || !totals || !P || !grad_part || !scratch || world < 1) { set_error("kl_gradient_targets: null pointer"); return -1; }
  TGArgs a{};
  a.k = kd; a.states = states; a.H = (int)H; a.Hp = H <= 32 ? 32 : 64;
  a.packed = packed; a.N = N; a.ld = ld; a.v = v; a.totals = totals; a.world = world; a.floor = floor;
  a.P = P; a.K = (int)K; a.R = (int)K * (kd.D + 1); a.p_stride = p_stride;
  a.part = (float*)scratch;
  a.klpart = a.part + (size_t)sm_count() * 128 * a.Hp;
  a.grad_out = grad_part; a.kl_out = kl_part; a.fault = fault;
  cudaStream_t st = (cudaStream_t)stream;
  switch (kd.D) {
    case 1: return launch_targets<1>(a, st);
    case 2: return launch_targets<2>(a, st);
    case 3: return launch_targets<3>(a, st);
    case 4: return launch_targets<4>(a, st);
    case 5: return launch_targets<5>(a, st);
    case 6: return launch_targets<6>(a, st);
    default: set_error("kl_gradient_targets: D=%d not instantiated (1..6)", kd.D); return -2;
  }
}

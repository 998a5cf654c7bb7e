// History footprint + spread of a planner step (klerg.py:470-475 and :496) with the squared distances on the
// tensor cores.
//
// The exponent of the Gaussian pair kernel in the expanded form is a bilinear form of D + 2 <= 8 terms,
//     e_ij = |sc_i|^2 + |xc_j|^2 - 2 xc_j . sc_i = a_i . b_j,
//     a_i = (sc_i[0..D), 0.., |sc_i|^2, 1),   b_j = (-2 xc_j[0..D), 0.., 1, |xc_j|^2)
// (sc = scaled sample - c, xc = scaled state - c, c = centre of the states' bounding box), i.e. exactly ONE K = 8
// step of a tf32 MMA.  With the 3xTF32 split (a_lo b_hi + a_hi b_lo + a_hi b_hi, both halves rounded to tf32, fp32
// accumulation) the result has fp32-grade accuracy as long as the states stay within the radius of the expanded form around c (the same bound the
// CUDA-core kernels use, KernelDev::x_r2); otherwise the launch falls back to footprint_kernel (gated on the device).
//
// What is left for the CUDA cores per pair is what no tensor core can do: min, ex2, add - the pass becomes
// MUFU-bound (1 exp per pair) instead of FP32-pipe-bound (D + 4 lane-ops per pair).
//
// One CTA per SM, 128 samples per tile (one sample per TMEM lane), all T state rows streamed past it in chunks of 128
// (the N of the MMA): D[128 samples][128 states] accumulators, three of them in flight.
//   warps 0-3, 4-7, 8-11  epilogue warpgroups, one accumulator buffer each (tcgen05.ld -> min / ex2 / add per
//                    sample); warps 0-3 also write the tile's A rows {hi | lo} into TMEM (tcgen05.st)
//   warp 12          streams the packed B chunks (8 KB: {hi | lo} x 2 K-halves x 128 rows x 16 B, K-major, no
//                    swizzle - the layout of klerg_targets_grad.cu) global -> shared with TMA bulk copies
//   warp 13          issues the MMAs (TS form: A from TMEM, B from shared memory), owns the TMEM allocation
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

constexpr int FT_ROWS = 128;                         // state rows per chunk (MMA N)
constexpr int FT_CHUNK_BYTES = 2 * 2 * FT_ROWS * 16; // {hi, lo} x {K 0-3, K 4-7} x rows x 16 B = 8 KB
constexpr int FT_STAGES = 8;                         // 64 KB ring
constexpr int FT_NWG = 3;                            // epilogue warpgroups = accumulator buffers in flight
constexpr int FT_THREADS = 32 * (4 * FT_NWG + 2);
constexpr int FT_HDR = 256;                          // scratch header: centre[8] floats, ok flag
constexpr int FT_ACOL = 128 * FT_NWG;                // TMEM columns: FT_NWG accumulators of 128, then A hi [8] | lo [8]
static_assert(FT_ACOL + 16 <= 512, "TMEM has 512 columns");

struct FTHeader {
  float ctr[8];
  int ok;       // 1: every state lies within the expanded form's radius of the centre -> the tensor-core pass runs
  int pad;
  unsigned long long next_tile;  // tiles are handed out on demand (zeroed by ft_centre_kernel before every pass)
  int pad2[4];
  double stats[4];               // {sum, max, min, #nan} of out_sum (klerg_vector_stats layout) -> totals
};
static_assert(sizeof(FTHeader) <= 256, "scratch header");
constexpr int FT_TSLOTS = 8;  // ring of tile announcements

struct FTArgs {
  KernelDev k;
  const float* states;
  long long T, T_sum;
  const float* packed;
  long long N, ld;
  float* out_sum;
  float* out_max;
  double* totals;
  void* ws;
  unsigned char* scratch;
  int nch_sum, nch_all;
  long long ntiles;
};

// ---- centre of the states' bounding box (scaled coordinates) and the radius test ------------------------------
__global__ void __launch_bounds__(256) ft_centre_kernel(const FTArgs a) {
  __shared__ float s_mn[8][8], s_mx[8][8], s_r2[8];
  FTHeader* h = reinterpret_cast<FTHeader*>(a.scratch);
  const int D = a.k.D, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mn[8], mx[8];
  for (int d = 0; d < 8; ++d) { mn[d] = INFINITY; mx[d] = -INFINITY; }
  for (long long j = threadIdx.x; j < a.T; j += blockDim.x)
    for (int d = 0; d < D; ++d) {
      const float x = a.states[j * a.k.S + a.k.explr[d]] * a.k.a[d];
      mn[d] = fminf(mn[d], x);
      mx[d] = fmaxf(mx[d], x);
    }
  for (int d = 0; d < D; ++d) {
    for (int o = 16; o; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    if (lane == 0) { s_mn[warp][d] = mn[d]; s_mx[warp][d] = mx[d]; }
  }
  __syncthreads();
  float c[8];
  for (int d = 0; d < 8; ++d) {
    c[d] = 0.f;
    if (d < D) {
      float lo = s_mn[0][d], hi = s_mx[0][d];
      for (int w = 1; w < 8; ++w) { lo = fminf(lo, s_mn[w][d]); hi = fmaxf(hi, s_mx[w][d]); }
      c[d] = 0.5f * (lo + hi);
    }
  }
  float r2 = 0.f;
  for (long long j = threadIdx.x; j < a.T; j += blockDim.x) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) {
      const float xc = a.states[j * a.k.S + a.k.explr[d]] * a.k.a[d] - c[d];
      s = fmaf(xc, xc, s);
    }
    r2 = fmaxf(r2, s);
  }
  for (int o = 16; o; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
  if (lane == 0) s_r2[warp] = r2;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) r2 = fmaxf(r2, s_r2[w]);
    for (int d = 0; d < 8; ++d) h->ctr[d] = c[d];
    h->ok = (r2 <= a.k.x_r2) ? 1 : 0;  // NaN rows fail the test too
    h->next_tile = 0ull;
  }
}

// ---- B chunks: b_j = (-2 xc_j, 0.., 1, |xc_j|^2) split into tf32 hi / lo, in the blocked K-major layout ---------
__global__ void __launch_bounds__(FT_ROWS) ft_pack_states_kernel(const FTArgs a) {
  const FTHeader* h = reinterpret_cast<const FTHeader*>(a.scratch);
  if (!h->ok) return;
  const int c = blockIdx.x, r = threadIdx.x, D = a.k.D;
  long long j;
  bool valid;
  if (c < a.nch_sum) {
    j = (long long)c * FT_ROWS + r;
    valid = j < a.T_sum;
  } else {
    j = a.T_sum + (long long)(c - a.nch_sum) * FT_ROWS + r;
    valid = j < a.T;
  }
  float b[8];
  for (int d = 0; d < 8; ++d) b[d] = 0.f;
  b[6] = 1.f;
  b[7] = 1e30f;  // rows beyond the list: e = 1e30 -> psi = 0, never the minimum
  if (valid) {
    float n = 0.f;
    for (int d = 0; d < D; ++d) {
      const float xc = a.states[j * a.k.S + a.k.explr[d]] * a.k.a[d] - h->ctr[d];
      b[d] = -2.f * xc;
      n = fmaf(xc, xc, n);
    }
    b[7] = n;
  }
  // Both halves are ROUNDED to tf32 (cvt.rna): truncating them - a plain mask for hi, the tensor core's own reading of
  // the low 13 bits for lo - doubles each half's error and quadruples the dropped lo * lo term; at the edge of the
  // radius that is the difference between 1.5e-4 and 4e-5 on psi (tests/test_tc_distance_model.py).
  uint32_t hi[8], lo[8];
  for (int d = 0; d < 8; ++d) {
    hi[d] = rna_tf32(b[d]);
    lo[d] = rna_tf32(b[d] - __uint_as_float(hi[d]));
  }
  unsigned char* base = a.scratch + FT_HDR + (size_t)c * FT_CHUNK_BYTES + (size_t)r * 16;
  *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + FT_ROWS * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  *reinterpret_cast<uint4*>(base + 2 * FT_ROWS * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  *reinterpret_cast<uint4*>(base + 3 * FT_ROWS * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}

// tcgen05.wait::ld that also "touches" the 32 destination registers, so that the compiler cannot move their uses
// above the wait (the loads are asynchronous until then)
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// Measured and not kept: taking every 3rd / 4th exponential off the MUFU pipe with a degree-4 polynomial on the FMA
// pipe (the FlashAttention-4 trick) changed the pass by < 4 % (278 / 286 against 288 ms with the switch compiled in):
// ncu shows the XU pipe at 81 % with the MIO queue as the top stall - the practical ceiling of MUFU-fed code here.
// eight independent sum chains and two min chains per thread: the dependent FADD / FMNMX latencies stay off the
// critical path of the three warps that share a scheduler
template <bool SUMMED>
__device__ __forceinline__ void ft_consume(const uint32_t (&r)[32], float (&s)[8], float (&emin)[2]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float e = __uint_as_float(r[j]);
    emin[(j >> 1) & 1] = fminf(emin[(j >> 1) & 1], e);
    if (SUMMED) s[j & 7] += ex2_neg(e);
  }
}

__global__ void __launch_bounds__(FT_THREADS, 1) footprint_tc_kernel(const FTArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FTHeader* hdr = reinterpret_cast<const FTHeader*>(a.scratch);
  if (!hdr->ok) return;  // the fallback launch behind this one does the pass
  unsigned char* ring = smem;
  unsigned long long* b_full = reinterpret_cast<unsigned long long*>(smem + (size_t)FT_STAGES * FT_CHUNK_BYTES);
  unsigned long long* b_empty = b_full + FT_STAGES;
  unsigned long long* acc_full = b_empty + FT_STAGES;  // [FT_NWG]
  unsigned long long* acc_empty = acc_full + FT_NWG;   // [FT_NWG]
  unsigned long long* a_full = acc_empty + FT_NWG;     // [1]
  unsigned long long* tile_ready = a_full + 1;         // [FT_TSLOTS]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tile_ready + FT_TSLOTS);
  volatile unsigned* s_abort = reinterpret_cast<volatile unsigned*>(s_tmem + 1);
  volatile int* s_tile = reinterpret_cast<volatile int*>(s_tmem + 4);  // [FT_TSLOTS] announced tile (or -1: no more)
  float* s_x = reinterpret_cast<float*>(s_tmem + 4 + FT_TSLOTS);  // [FT_NWG - 1][128][2]: the other warpgroups' {sum, min} of the tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = a.nch_all;
  if (threadIdx.x == 0) {
    for (int s = 0; s < FT_STAGES; ++s) {
      bar_init(&b_full[s], 1);
      bar_init(&b_empty[s], 1);
    }
    for (int g = 0; g < FT_NWG; ++g) {
      bar_init(&acc_full[g], 1);
      bar_init(&acc_empty[g], 4);
    }
    bar_init(a_full, 128);
    for (int q = 0; q < FT_TSLOTS; ++q) bar_init(&tile_ready[q], 1);
    *s_abort = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 * FT_NWG + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  // Tiles (128 samples) are handed out on demand: the TMA thread - the role that runs furthest ahead - draws the next
  // tile number from a global counter and announces it (or -1) to the other roles through a ring of FT_TSLOTS slots,
  // each guarded by an mbarrier.  An SM that is slowed down (it may host the side-stream draw of the next step,
  // engine.UniformPrefetch) simply takes fewer tiles.  Slot reuse is safe by construction: the TMA thread runs at most
  // FT_STAGES + FT_NWG chunks ahead of the slowest epilogue warp, i.e. < FT_TSLOTS tiles with >= 2 chunks per tile.

  if (warp < 4 * FT_NWG) {
    // ================= epilogue warpgroups =================
    const int g = warp >> 2;                       // accumulator buffer of this warpgroup
    const int row = (warp & 3) * 32 + lane;        // sample of the tile = TMEM lane
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int D = a.k.D;
    float ctr[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) ctr[d] = hdr->ctr[d];
    long long gc = 0;  // chunks of this CTA so far (all tiles)
    for (int it = 0;; ++it) {
      bar_wait(&tile_ready[it & (FT_TSLOTS - 1)], (unsigned)((it / FT_TSLOTS) & 1), s_abort);
      const long long tile = s_tile[it & (FT_TSLOTS - 1)];
      if (tile < 0 || *s_abort) break;
      const long long i = tile * 128 + row;
      if (g == 0) {
        // A rows of the tile: a_i = (sc_i, 0.., |sc_i|^2, 1) as tf32 hi | lo
        float av[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) av[d] = 0.f;
        float n = 0.f;
        if (i < a.N) {
          for (int d = 0; d < D; ++d) {
            const float sc = __ldg(a.packed + (long long)d * a.ld + i) - ctr[d];
            av[d] = sc;
            n = fmaf(sc, sc, n);
          }
        }
        av[6] = n;
        av[7] = 1.f;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          hi[d] = rna_tf32(av[d]);
          lo[d] = rna_tf32(av[d] - __uint_as_float(hi[d]));
        }
        tmem_st8(trow + FT_ACOL, hi);
        tmem_st8(trow + FT_ACOL + 8, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        bar_arrive(a_full);
      }
      float total = 0.f, emin2[2] = {INFINITY, INFINITY};
      for (int c = 0; c < nch; ++c, ++gc) {
        if ((int)(gc % FT_NWG) != g) continue;
        bar_wait(&acc_full[g], (unsigned)((gc / FT_NWG) & 1), s_abort);
        tc_fence_after();
        const uint32_t acc = trow + (uint32_t)(g * 128);
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint32_t ra[32], rb[32];
        const bool summed = c < a.nch_sum;
#define FT_CONSUME(R)                       \
  if (summed) ft_consume<true>(R, s, emin2); \
  else ft_consume<false>(R, s, emin2)
        tmem_ld32(acc, ra);
        tmem_ld_wait32(ra);
        tmem_ld32(acc + 32, rb);
        FT_CONSUME(ra);
        tmem_ld_wait32(rb);
        tmem_ld32(acc + 64, ra);
        FT_CONSUME(rb);
        tmem_ld_wait32(ra);
        tmem_ld32(acc + 96, rb);
        FT_CONSUME(ra);
        tmem_ld_wait32(rb);
        // the accumulator is in registers: hand the buffer back before the last quarter is consumed
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(&acc_empty[g]);
        FT_CONSUME(rb);
        total += ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));  // per-chunk partial: the rounding error grows with sqrt(chunks)
      }
      // combine the two warpgroups; all chunks of the tile are consumed -> every MMA that read A has completed
      float emin = fminf(emin2[0], emin2[1]);
      if (g > 0) {
        s_x[((g - 1) * 128 + row) * 2] = total;
        s_x[((g - 1) * 128 + row) * 2 + 1] = emin;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * FT_NWG) : "memory");
      if (g == 0) {
#pragma unroll
        for (int o = 0; o < FT_NWG - 1; ++o) {  // fixed order
          total += s_x[(o * 128 + row) * 2];
          emin = fminf(emin, s_x[(o * 128 + row) * 2 + 1]);
        }
        if (i < a.N) {
          a.out_sum[i] = total * a.k.inv_nu;
          a.out_max[i] = ex2_neg(emin) * a.k.inv_nu;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * FT_NWG) : "memory");  // s_x may be rewritten, A may be overwritten
    }
  } else if (warp == 4 * FT_NWG) {
    // ================= B chunks: TMA bulk copies into the ring =================
    if (lane == 0) {
      const unsigned char* src = a.scratch + FT_HDR;
      unsigned long long* next = &reinterpret_cast<FTHeader*>(a.scratch)->next_tile;
      long long gc = 0;
      for (int it = 0;; ++it) {
        const unsigned long long t = atomicAdd(next, 1ull);
        s_tile[it & (FT_TSLOTS - 1)] = t < (unsigned long long)a.ntiles ? (int)t : -1;
        bar_arrive(&tile_ready[it & (FT_TSLOTS - 1)]);
        if (t >= (unsigned long long)a.ntiles || *s_abort) break;
        for (int c = 0; c < nch; ++c, ++gc) {
          const int s = (int)(gc % FT_STAGES);
          const unsigned k = (unsigned)(gc / FT_STAGES);
          bar_wait(&b_empty[s], (k & 1u) ^ 1u, s_abort);
          bar_expect_tx(&b_full[s], FT_CHUNK_BYTES);
          bulk_g2s(ring + (size_t)s * FT_CHUNK_BYTES, src + (size_t)c * FT_CHUNK_BYTES, FT_CHUNK_BYTES, &b_full[s]);
        }
      }
    }
  } else {
    // ================= MMA issue =================
    if (lane == 0) {
      const uint32_t idesc = instr_desc_tf32(128, FT_ROWS);
      const uint32_t a_hi = tmem + FT_ACOL, a_lo = a_hi + 8;
      long long gc = 0;
      for (int it = 0;; ++it) {
        bar_wait(&tile_ready[it & (FT_TSLOTS - 1)], (unsigned)((it / FT_TSLOTS) & 1), s_abort);
        if (s_tile[it & (FT_TSLOTS - 1)] < 0 || *s_abort) break;
        bar_wait(a_full, (unsigned)(it & 1), s_abort);
        tc_fence_after();
        for (int c = 0; c < nch; ++c, ++gc) {
          const int s = (int)(gc % FT_STAGES), buf = (int)(gc % FT_NWG);
          const unsigned k = (unsigned)(gc / FT_STAGES);
          bar_wait(&b_full[s], k & 1u, s_abort);
          bar_wait(&acc_empty[buf], (unsigned)((gc / FT_NWG) & 1) ^ 1u, s_abort);
          tc_fence_after();
          const uint32_t b0 = smem_addr(ring + (size_t)s * FT_CHUNK_BYTES);
          const uint64_t db_hi = smem_desc(b0, FT_ROWS * 16, 128), db_lo = smem_desc(b0 + 2 * FT_ROWS * 16, FT_ROWS * 16, 128);
          const uint32_t d = tmem + (uint32_t)(buf * 128);
          tc_mma_tf32_ts(d, a_lo, db_hi, idesc, 0u);
          tc_mma_tf32_ts(d, a_hi, db_lo, idesc, 1u);
          tc_mma_tf32_ts(d, a_hi, db_hi, idesc, 1u);
          tc_commit(&b_empty[s]);
          tc_commit(&acc_full[buf]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4 * FT_NWG + 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
  if (threadIdx.x == 0 && *s_abort) atomicExch(ws_fused_ctrl(a.ws) + 5, 1u);  // the sticky fault word of the workspace
}

// totals[0..1] = {sum, max} of out_sum from the stats a klerg_vector_stats pass left in the header (fixed summation
// order: the totals do not depend on which SM took which tile)
__global__ void ft_totals_kernel(const unsigned char* scratch, double* totals) {
  const FTHeader* h = reinterpret_cast<const FTHeader*>(scratch);
  if (!h->ok) return;  // the fallback pass writes its own
  totals[0] = h->stats[0];
  totals[1] = h->stats[1];
}

}  // namespace

// declared in klerg_pairwise.cu: the CUDA-core pass, skipped on the device when *gate != 0
int launch_footprint_sum_max_gated(const KernelDev& kd, const float* states, int64_t T, int64_t T_sum, const float* packed,
                                   int64_t N, int64_t ld, float* out_sum, float* out_max, double* totals, void* workspace,
                                   const int* gate, cudaStream_t stream);
}  // namespace klerg

using namespace klerg;

extern "C" int64_t klerg_footprint_tc_scratch_bytes(int64_t T) {
  if (T < 0) return -1;
  return (int64_t)FT_HDR + ((T + FT_ROWS - 1) / FT_ROWS + 2) * (int64_t)FT_CHUNK_BYTES;
}

extern "C" int klerg_footprint_sum_max_tc(const klerg_kernel_spec* k, const float* states, int64_t T, int64_t T_sum,
                                          const float* packed, int64_t N, int64_t ld, float* out_sum, float* out_max,
                                          double* totals, void* workspace, void* scratch, int64_t scratch_bytes,
                                          void* stream) {
  KernelDev kd;
  if (!make_kernel_dev(k, kd)) return -1;
  if (N < 1 || T < 1 || T_sum < 0 || T_sum > T || ld < N || (ld & 3)) { set_error("footprint_sum_max_tc: bad sizes"); return -1; }
  if (!workspace || !totals || !out_sum || !out_max || !scratch) { set_error("footprint_sum_max_tc: null output/workspace/scratch"); return -1; }
  if (kd.D > 6) { set_error("footprint_sum_max_tc: D + 2 must fit one K = 8 step (D <= 6)"); return -1; }
  if (scratch_bytes < klerg_footprint_tc_scratch_bytes(T) || ((uintptr_t)scratch & 127)) {
    set_error("footprint_sum_max_tc: scratch too small or not 128-byte aligned");
    return -1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  FTArgs a{};
  a.k = kd; a.states = states; a.T = T; a.T_sum = T_sum; a.packed = packed; a.N = N; a.ld = ld;
  a.out_sum = out_sum; a.out_max = out_max; a.totals = totals; a.ws = workspace; a.scratch = (unsigned char*)scratch;
  a.nch_sum = (int)((T_sum + FT_ROWS - 1) / FT_ROWS);
  a.nch_all = a.nch_sum + (int)((T - T_sum + FT_ROWS - 1) / FT_ROWS);
  a.ntiles = (N + 127) / 128;
  if (a.nch_all < 2) { set_error("footprint_sum_max_tc: at least two 128-row chunks of states (T > 128) are required"); return -1; }
  ft_centre_kernel<<<1, 256, 0, st>>>(a);
  if (int rc = check_launch("ft_centre_kernel")) return rc;
  ft_pack_states_kernel<<<(unsigned)a.nch_all, FT_ROWS, 0, st>>>(a);
  if (int rc = check_launch("ft_pack_states_kernel")) return rc;
  size_t smem = (size_t)FT_STAGES * FT_CHUNK_BYTES + 8 * (2 * FT_STAGES + 2 * FT_NWG + 1 + FT_TSLOTS) + 16 + 4 * FT_TSLOTS +
                sizeof(float) * 256 * (FT_NWG - 1) + 64;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM: each allocates all 512 TMEM columns
  static bool raised = false;
  if (!raised) {
    cudaError_t e = cudaFuncSetAttribute(footprint_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("footprint_sum_max_tc: shared memory attribute: %s", cudaGetErrorString(e)); return -4; }
    raised = true;
  }
  long long grid = sm_count();
  if (grid > a.ntiles) grid = a.ntiles;
  footprint_tc_kernel<<<(unsigned)grid, FT_THREADS, smem, st>>>(a);
  if (int rc = check_launch("footprint_tc_kernel")) return rc;
  if (int rc = klerg_vector_stats(out_sum, N, reinterpret_cast<FTHeader*>(scratch)->stats, workspace, stream)) return rc;
  ft_totals_kernel<<<1, 1, 0, st>>>((const unsigned char*)scratch, totals);
  if (int rc = check_launch("ft_totals_kernel")) return rc;
  // states outside the expanded form's radius: the CUDA-core pass (returns at once when the header says ok)
  return launch_footprint_sum_max_gated(kd, states, T, T_sum, packed, N, ld, out_sum, out_max, totals, workspace,
                                        &reinterpret_cast<const FTHeader*>(scratch)->ok, st);
}

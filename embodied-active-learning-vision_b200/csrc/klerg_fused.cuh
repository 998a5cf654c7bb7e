// Fused KL-ergodic evals: ONE launch per eval (sm_100a).
//
//   eval_grad_kernel : Robot.forward + footprint + renormalize + importance ratio +
//                      kldiv_grad_vec for all H steps + Robot.backward   (klerg.py:409-450, 505-523)
//                      -> du, djdlam, u*, dgdx, KL cost of the plan
//   eval_cost_kernel : Robot.get_cost for G <= 8 candidate control sequences (klerg.py:686-710;
//                      the <= 5 line-search windows of klerg.py:712-751 are one launch)
//
// Every CTA redoes the tiny rollout (states only) in shared memory, owns a contiguous slice of the workspace
// samples, and the CTAs - of this GPU and, with a sample-sharded workspace, of every peer GPU - meet twice:
//   (1) after the forward pair pass: all-reduce of {sum, max} of q = q_base + q_iter (renormalize needs both
//       before the importance ratio exists).  One hop: every CTA stores its pair as tagged words into the mailbox
//       of every rank (plain stores, NVLink for the peers) and polls its own GPU's mailbox until all
//       world x CTAs pairs are there (klerg_ll.cuh).  No atomics, no leader, no second broadcast.
//   (2) after the gradient pass: every CTA stores its H*D fp32 partials as tagged words; warps spread over the
//       grid poll one gradient entry each, add the CTA partials in a fixed order and store the sum into every
//       rank's mailbox; ONE CTA (the "finisher") polls those sums, adds them in rank order and runs the adjoint
//       sweep.  All other CTAs leave right after their entry sums.
// The launches carry the programmatic-dependent-launch attribute and trigger their dependents right after
// meeting (1): the CTAs of the next eval in the stream start on SMs as the CTAs of this one leave, so the
// finisher's tail (gather + adjoint, ~5 us) and the launch latency overlap the next eval's rollout / forward
// pass.  Full grids leave one SM free so that the next eval never has to wait for the finisher's SM.  By
// default the next eval still waits (griddepcontrol.wait) for this one to complete before it reads its
// inputs; KLERG_OPT_EVAL_OVERLAP declares consecutive evals independent and drops that wait.
#pragma once
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "klerg_common.cuh"
#include "klerg_dyn.cuh"
#include "klerg_ll.cuh"
#include "klerg_pair.cuh"

namespace klerg {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
}
// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier -------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// one contiguous run of `bytes` (multiple of 16, both sides 16-byte aligned) global -> shared
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Phase stamps (clock64 at the phase boundaries of eval_grad_kernel) are compiled in with -DKLERG_STAMPS.
#ifdef KLERG_STAMPS
#define KLERG_STAMP_DECL long long stamp[16]
#define KLERG_STAMP(i) stamp[i] = clock64()
// per-CTA wall-clock stamps (ns, comparable across SMs) in the debug region of the local mailbox
#define KLERG_CTA_STAMP(me, vblk, i)                                                   \
  if (threadIdx.x == 0) {                                                               \
    unsigned long long t_;                                                              \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                              \
    ((unsigned long long*)((char*)(me) + MB_OFF_DBG))[(size_t)(vblk) * 8 + (i)] = t_;   \
  }
#else
#define KLERG_STAMP_DECL
#define KLERG_STAMP(i)
#define KLERG_CTA_STAMP(me, vblk, i)
#endif

struct EvalArgs {
  KernelDev k;
  DynDev d;
  BarDev bar;
  AdjParams ap;
  Peers peers;
  int independent;       // 1: inputs do not depend on the previous launch in the stream (no griddepcontrol.wait)
  // Device-side gate (klerg_plan_optimize): a word written by an earlier kernel of the stream.  Cost eval: the number
  // of candidates actually evaluated (<= G; 0: the launch does nothing); gradient eval: 0 = the launch does nothing.
  // The same value on every rank, so a skipped launch is skipped everywhere (no exchange is left half-done).
  const int* gate;
  // inputs
  const float* x0;       // [S]
  const float* R0;       // [9] or NULL
  const float* u;        // [G][H][A]
  int G, H;
  const float* packed;   // [D][ld] scaled samples of this rank
  int64_t N, ld;
  const float* q_base;   // [N] or NULL
  const float* p;        // [K][p_stride] target densities (K = 1: [N])
  int K;                 // belief targets sharing one workspace / trajectory (gradient eval)
  int64_t p_stride;
  const double* p_stats; // [K] = sum p_k over all ranks
  float floor;
  // scratch
  float* v;              // [G][ld]
  void* ws;
  // gradient-mode schedule
  int nchr, nsub, rounds;  // state chunks per round, sample sub-streams, rounds
  int nwide;               // mixed schedule: the last `nwide` warps own WT+1 states, the others WT
  int ts;                  // samples staged per tile (multiple of 64)
  // outputs
  float* traj;           // [G][H+1][S] or NULL
  double* totals;        // [G][2] {sum, max} of q_base + q_iter over all ranks, or NULL
  float* cost;           // [G]
  float* dgdx;           // [H][S]
  float* du;             // [H][A]
  float* djdlam;         // [H]
  float* u_star;         // [H][A]
  float* fault_out;      // [1] or NULL: 0 / 1 copy of the sticky fault word next to the outputs the host reads
  double* kl_out;        // [2] {sum p(log p - log c), sum c} over all ranks, or NULL
};

extern const int* g_eval_gate;  // EvalArgs::gate of the evals launched while it is set (klerg_plan.cu)

// process-wide switches of the fused evals (klerg_set_option)
struct FusedOptions {
  int overlap;      // consecutive evals are independent: no griddepcontrol.wait
  int grid_limit;   // > 0: at most this many CTAs per launch (tests: two emulated ranks share one GPU)
  int pdl;          // launch with the programmatic-stream-serialization attribute
  int coop_probe;   // -1 unknown, 0 / 1: the cooperative attribute may be combined with it
  int mixed_warps;  // 12 = balanced 12-warp schedule for D >= 5 where the horizon allows (A/B; measured slower than the default)
};
extern FusedOptions g_fused_opt;

// World-2 emulation on ONE GPU (tests): the two ranks' launches are recorded and run as one cooperative grid
// whose first half acts as rank 0 and second half as rank 1 (separate launches that wait on one another are
// not guaranteed to be co-resident).
struct EmuState {
  int active;
  int have[2];
  EvalArgs args[2];
  int (*launch)(const EvalArgs&, const EvalArgs&, int, int, size_t, cudaStream_t);
  int nblk, nthreads;
  size_t smem;
};
extern EmuState g_emu;

// block reduction of NQ doubles (fixed order); result valid in thread 0
template <int NQ>
__device__ __forceinline__ void block_reduce(const int (&kind)[NQ], double (&val)[NQ], double* sh_red /* [32*NQ] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const double v = warp_reduce(kind[q], val[q]);
    if (lane == 0) sh_red[warp * NQ + q] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      double v = lane < nwarp ? sh_red[lane * NQ + q] : red_identity(kind[q]);
      v = warp_reduce(kind[q], v);
      if (lane == 0) val[q] = v;
    }
  }
}

// Block reduction of G pairs {a_g, b_g} in one go (two barriers in total instead of two per candidate):
// kind_b = RED_MAX or RED_SUM for the second member; results for all g valid in thread 0.
__device__ __forceinline__ void block_reduce_pairs(int G, int kind_b, double (&va)[FUSED_MAXG], double (&vb)[FUSED_MAXG],
                                                   double* sh_red /* [32 * 2 * FUSED_MAXG] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  __syncthreads();
#pragma unroll
  for (int g = 0; g < FUSED_MAXG; ++g) {
    if (g < G) {
      const double x = warp_reduce(RED_SUM, va[g]);
      const double y = warp_reduce(kind_b, vb[g]);
      if (lane == 0) {
        sh_red[(warp * FUSED_MAXG + g) * 2 + 0] = x;
        sh_red[(warp * FUSED_MAXG + g) * 2 + 1] = y;
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int g = 0; g < FUSED_MAXG; ++g) {
      if (g < G) {
        double x = lane < nwarp ? sh_red[(lane * FUSED_MAXG + g) * 2 + 0] : 0.0;
        double y = lane < nwarp ? sh_red[(lane * FUSED_MAXG + g) * 2 + 1] : red_identity(kind_b);
        x = warp_reduce(RED_SUM, x);
        y = warp_reduce(kind_b, y);
        if (lane == 0) {
          va[g] = x;
          vb[g] = y;
        }
      }
    }
  }
}

// contiguous sample slice of CTA vblk of vnblk: [lo, hi) with lo % 8 == 0, hi <= ld
__device__ __forceinline__ void cta_slice(int64_t N, int64_t ld, int vblk, int vnblk, int64_t& lo, int64_t& hi) {
  int64_t per = (N + vnblk - 1) / vnblk;
  per = (per + 7) & ~(int64_t)7;
  lo = (int64_t)vblk * per;
  hi = lo + per;
  if (hi > ld) hi = ld;
  if (lo > ld) lo = ld;
}

// State rows of one eval in shared memory, in both forms (klerg_pair.cuh): duplicated rows for the difference form,
// {-2 xc, |xc|^2} rows around the centre `ctr` for the expanded form; `xform` says which one the pair passes use.
struct StateRows {
  const u64* x2;     // [T][DP]
  const float* rx;   // [T][NF]
  const float* ctr;  // [D]
  bool xform;
};

// Forward pair pass of one trajectory (T state rows) over this CTA's slice:
// v[i] = q_base[i] + inv_nu * sum_t psi; returns the slice's {sum, max} of v over i < N.
template <int D, int P>
__device__ __forceinline__ void forward_slice_p(const EvalArgs& a, const StateRows& sr, int T, float* v_out, int64_t lo,
                                                int64_t hi, double& tsum, double& tmax) {
  constexpr int SPT = 2 * P;
  u64 c2[D];
#pragma unroll
  for (int d = 0; d < D; ++d) c2[d] = pack2(sr.ctr[d], sr.ctr[d]);
  // this thread's next samples are requested before the pair math of the current ones starts (a thread-iteration
  // is only T state rows long, and the load latency would otherwise be exposed once per iteration); they stay in
  // the registers the loads return them in until they become current, so that nothing waits for them early
  typedef typename std::conditional<P == 2, float4, float2>::type vec_t;
  auto load = [&](int64_t i0, vec_t (&raw)[D], float (&qb)[SPT]) {
#pragma unroll
    for (int d = 0; d < D; ++d) raw[d] = __ldg(reinterpret_cast<const vec_t*>(a.packed + (int64_t)d * a.ld + i0));
#pragma unroll
    for (int q = 0; q < SPT; ++q) qb[q] = (a.q_base && i0 + q < a.N) ? __ldg(a.q_base + i0 + q) : 0.f;
  };
  const int64_t stride = (int64_t)blockDim.x * SPT;
  int64_t i0 = lo + (int64_t)threadIdx.x * SPT;
  vec_t raw[D], raw_n[D];
  float qb[SPT], qbn[SPT];
  if (i0 < hi) load(i0, raw, qb);
  for (; i0 < hi; i0 += stride) {
    u64 s2[D][P], acc[P];
    float emin[SPT];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if constexpr (P == 2) {
        s2[d][0] = pack2(raw[d].x, raw[d].y);
        s2[d][1] = pack2(raw[d].z, raw[d].w);
      } else {
        s2[d][0] = pack2(raw[d].x, raw[d].y);
      }
    }
    const bool more = i0 + stride < hi;
#pragma unroll
    for (int q = 0; q < P; ++q) acc[q] = pack2(0.f, 0.f);
    u64 s2n[P];
    if (sr.xform) {
#pragma unroll
      for (int q = 0; q < P; ++q) {
        s2[0][q] = sub2(s2[0][q], c2[0]);
        s2n[q] = mul2(s2[0][q], s2[0][q]);
#pragma unroll
        for (int d = 1; d < D; ++d) {
          s2[d][q] = sub2(s2[d][q], c2[d]);
          s2n[q] = fma2(s2[d][q], s2[d][q], s2n[q]);
        }
      }
    }
    // The next samples are requested only now, after the current ones have been consumed: a consumer of loaded data
    // waits for every load outstanding on its scoreboard, including ones issued after the load it depends on.
    asm volatile("" ::: "memory");
    if (more) load(i0 + stride, raw_n, qbn);
    asm volatile("" ::: "memory");
    if (sr.xform)
      pair_forward_x<D, P, 0>(sr.rx, 0, T, s2, s2n, acc, emin);
    else
      pair_forward<D, P, 0>(sr.x2, 0, T, s2, acc, emin);
    float o[SPT];
#pragma unroll
    for (int q = 0; q < P; ++q) unpack2(acc[q], o[2 * q], o[2 * q + 1]);
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
      const int64_t i = i0 + q;
      float v = o[q] * a.k.inv_nu;
      if (i < a.N) {
        if (a.q_base) v += qb[q];
        tsum += (double)v;
        tmax = fmax(tmax, (double)v);
      }
      o[q] = v;
    }
    if constexpr (P == 2)
      *reinterpret_cast<float4*>(v_out + i0) = make_float4(o[0], o[1], o[2], o[3]);
    else
      *reinterpret_cast<float2*>(v_out + i0) = make_float2(o[0], o[1]);
    if (more) {
#pragma unroll
      for (int d = 0; d < D; ++d) raw[d] = raw_n[d];
#pragma unroll
      for (int q = 0; q < SPT; ++q) qb[q] = qbn[q];
    }
  }
}

// Samples per thread-iteration: 2 (one packed pair) or 4.  A slice of n samples costs ceil(n / (threads * 2P)) * P
// pair-iterations per thread; for slices of a few samples per thread the quantisation decides (e.g. 6757 samples on
// 512 threads: 4 iterations of two pairs = 8, or 7 iterations of one pair = 7).  A pair-iteration of the two-pair form
// is ~10 % cheaper (the state rows are read from shared memory once for both pairs: 3.19e12 against 2.86e12 pairs/s
// in the isolated D = 6 loop), so the one-pair form has to save more than that to be chosen.
__device__ __forceinline__ bool narrow_pairs(int64_t lo, int64_t hi) {
  const int64_t n = hi - lo, bd = blockDim.x;
  const int64_t it1 = (n + bd * 2 - 1) / (bd * 2), it2 = (n + bd * 4 - 1) / (bd * 4);
  return 10 * it1 < 18 * it2 || it1 <= 1;
}

template <int D>
__device__ __forceinline__ void forward_slice(const EvalArgs& a, const StateRows& sr, int T, float* v_out, int64_t lo,
                                              int64_t hi, double& tsum, double& tmax) {
  if (narrow_pairs(lo, hi))
    forward_slice_p<D, 1>(a, sr, T, v_out, lo, hi, tsum, tmax);
  else
    forward_slice_p<D, 2>(a, sr, T, v_out, lo, hi, tsum, tmax);
}

// Stage the state rows of G trajectories of T rows each in both forms.  Row t of trajectory g is at
// states[g * seg + t * S] (explored columns a.k.explr, scaled by a.k.a); centre = the middle row of trajectory 0.
// s_xs[G*T][D] (may be NULL) receives the centred scaled coordinates.  Returns (all threads) whether every row
// lies within the radius of the expanded form.  Ends with a barrier.
template <int D>
__device__ __forceinline__ bool stage_state_rows(const EvalArgs& a, const float* states, int S, int G, int T, size_t seg,
                                                 u64* s_x2, float* s_rx, float* s_xs, float* s_ctr) {
  constexpr int DP = Row2<D>::DP, NF = RowX<D>::NF;
  const int tid = threadIdx.x;
  float c[D];
#pragma unroll
  for (int d = 0; d < D; ++d) c[d] = states[(size_t)(T / 2) * S + a.k.explr[d]] * a.k.a[d];
  if (tid < D) s_ctr[tid] = states[(size_t)(T / 2) * S + a.k.explr[tid]] * a.k.a[tid];
  bool big = false;
  for (int r = tid; r < G * T; r += blockDim.x) {
    const int g = r / T, t = r - g * T;
    const float* row = states + (size_t)g * seg + (size_t)t * S;
    float x2n = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float x = row[a.k.explr[d]] * a.k.a[d];
      const float xc = x - c[d];
      if (s_xs) s_xs[r * D + d] = xc;
      s_x2[r * DP + d] = pack2(x, x);
      s_rx[r * NF + d] = -2.f * xc;
      x2n = fmaf(xc, xc, x2n);
    }
#pragma unroll
    for (int d = D; d < DP; ++d) s_x2[r * DP + d] = pack2(0.f, 0.f);
    s_rx[r * NF + D] = x2n;
#pragma unroll
    for (int d = D + 1; d < NF; ++d) s_rx[r * NF + d] = 0.f;
    big |= !(x2n <= a.k.x_r2);
  }
  return !__syncthreads_or(big);
}

// Forward pair pass of G candidate trajectories over this CTA's slice: samples are loaded once per thread and swept
// against every candidate (the per-candidate totals live in a small indexed array: two local-memory accesses per
// H pairs); v[g][i] to HBM; the CTA's {sum, max} per candidate end up in s_in[2g], s_in[2g + 1].
template <int D, int P>
__device__ __forceinline__ void forward_candidates(const EvalArgs& a, const StateRows& sr, int G, int H, int64_t lo, int64_t hi,
                                                   double* s_red, double* s_in, float* vbase) {
  constexpr int SPT = 2 * P, DP = Row2<D>::DP, NF = RowX<D>::NF;
  const int tid = threadIdx.x;
  double tsum[FUSED_MAXG], tmax[FUSED_MAXG];
  for (int g = 0; g < FUSED_MAXG; ++g) {
    tsum[g] = 0.0;
    tmax[g] = -INFINITY;
  }
  u64 c2[D];
#pragma unroll
  for (int d = 0; d < D; ++d) c2[d] = pack2(sr.ctr[d], sr.ctr[d]);
  for (int64_t i0 = lo + (int64_t)tid * SPT; i0 < hi; i0 += (int64_t)blockDim.x * SPT) {
    u64 s2[D][P], s2n[P];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if constexpr (P == 2) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(a.packed + (int64_t)d * a.ld + i0));
        s2[d][0] = pack2(s.x, s.y);
        s2[d][1] = pack2(s.z, s.w);
      } else {
        const float2 s = __ldg(reinterpret_cast<const float2*>(a.packed + (int64_t)d * a.ld + i0));
        s2[d][0] = pack2(s.x, s.y);
      }
    }
#pragma unroll
    for (int q = 0; q < P; ++q) s2n[q] = pack2(0.f, 0.f);
    if (sr.xform) {
#pragma unroll
      for (int q = 0; q < P; ++q) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          s2[d][q] = sub2(s2[d][q], c2[d]);
          s2n[q] = fma2(s2[d][q], s2[d][q], s2n[q]);
        }
      }
    }
    float qb[SPT];
#pragma unroll
    for (int q = 0; q < SPT; ++q) qb[q] = (a.q_base && i0 + q < a.N) ? a.q_base[i0 + q] : 0.f;
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      u64 acc[P];
      float emin[SPT], o[SPT];
#pragma unroll
      for (int q = 0; q < P; ++q) acc[q] = pack2(0.f, 0.f);
      if (sr.xform)
        pair_forward_x<D, P, 0>(sr.rx + (size_t)g * H * NF, 0, H, s2, s2n, acc, emin);
      else
        pair_forward<D, P, 0>(sr.x2 + (size_t)g * H * DP, 0, H, s2, acc, emin);
#pragma unroll
      for (int q = 0; q < P; ++q) unpack2(acc[q], o[2 * q], o[2 * q + 1]);
      double ts = tsum[g], tm = tmax[g];
#pragma unroll
      for (int q = 0; q < SPT; ++q) {
        float v = o[q] * a.k.inv_nu;
        if (i0 + q < a.N) {
          v += qb[q];
          ts += (double)v;
          tm = fmax(tm, (double)v);
        }
        o[q] = v;
      }
      tsum[g] = ts;
      tmax[g] = tm;
      float* vp = vbase + (size_t)g * a.ld + i0;
      if constexpr (P == 2)
        *reinterpret_cast<float4*>(vp) = make_float4(o[0], o[1], o[2], o[3]);
      else
        *reinterpret_cast<float2*>(vp) = make_float2(o[0], o[1]);
    }
  }
  double ra[FUSED_MAXG], rb[FUSED_MAXG];
#pragma unroll
  for (int g = 0; g < FUSED_MAXG; ++g) {
    ra[g] = tsum[g];
    rb[g] = tmax[g];
  }
  block_reduce_pairs(G, RED_MAX, ra, rb, s_red);
  if (tid == 0) {
#pragma unroll
    for (int g = 0; g < FUSED_MAXG; ++g)
      if (g < G) {
        s_in[2 * g] = ra[g];
        s_in[2 * g + 1] = rb[g];
      }
  }
}

// ---------------------------------------------------------------------------
// shared-memory carve-up
// ---------------------------------------------------------------------------
// Row stride (floats) of the staged sample tiles: a compile-time constant so that the D+2 row addresses of a
// tile are immediates off one base register (no address chain, fewer live registers in the pair loop).
#define TILE_ROWS(D) ((D) + 3)
// Sample tiles of the gradient pass: two buffers of TS_ROW samples; TMA fills one while the warps work on the other.
// (A deeper ring of smaller tiles without the per-tile CTA barrier was measured and is slower: the per-tile
// bookkeeping of every warp outweighs the barrier it removes.)
constexpr int TILE_NB = 2;
constexpr int TS_ROW = 2048;

struct SmemPlan {
  size_t u, traj, dbarr, P, rot, x2, rx, xs, tile, adj, part, red, misc, ll, total;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
constexpr size_t LL_SCRATCH_BYTES = sizeof(double) * (LL_BUF_VALS + 32);
constexpr size_t MISC_BYTES = 16 + sizeof(float) * (KLERG_MAX_S + 9) + 16 + 8 + 8 * 4 + 8;

template <int D>
__host__ __device__ inline SmemPlan plan_grad(int H, int S, int A, bool roll, int nwarps, int WT, int tsr) {
  SmemPlan p{};
  size_t o = 0;
  p.u = o;     o = align16(o + sizeof(float) * H * A);
  p.traj = o;  o = align16(o + sizeof(float) * (H + 1) * S);
  p.dbarr = o; o = align16(o + sizeof(float) * H * S);
  p.P = o;     o = align16(o + (roll ? sizeof(float) * H * A * A : 0));
  p.rot = o;   o = align16(o + (roll ? sizeof(float) * rollout_rot_floats(1, H) : 0));  // kept for the finisher
  p.x2 = o;    o = align16(o + sizeof(u64) * H * Row2<D>::DP);
  p.rx = o;    o = align16(o + sizeof(float) * H * RowX<D>::NF);
  p.xs = o;    o = align16(o + sizeof(float) * H * D);
  // ring of TMA buffers of rows s_0..s_{D-1}, v -> w, p (+ one row |sc|^2 computed in place); its second half
  // doubles as the staging area of meeting (1), before the tiles that live there are requested
  p.tile = o;  o = align16(o + sizeof(float) * (size_t)TILE_NB * TILE_ROWS(D) * tsr);
  // the finisher's adjoint: gathered sums, dgdx, scratch, KL staging
  p.adj = o;   o = align16(o + sizeof(double) * ((size_t)H * (D + 1) + 2) +
                           sizeof(float) * ((size_t)H * S + adjoint_scratch_floats(H, A)) + sizeof(double) * (2 * LL_MAXBLK + 40));
  p.part = o;  o = align16(o + sizeof(float) * (size_t)nwarps * WT * (D + 1));
  p.red = o;   o = align16(o + sizeof(double) * 32 * 4);
  p.misc = o;  o = align16(o + MISC_BYTES);
  p.total = o;
  return p;
}
static_assert(sizeof(float) * TILE_ROWS(1) * TS_ROW >= LL_SCRATCH_BYTES, "meeting staging must fit one tile buffer (D = 1)");

// ---------------------------------------------------------------------------
// gradient eval
// ---------------------------------------------------------------------------
// MIXED, LEFT > 0 (balanced): every warp owns exactly WT states and the <= LEFT states that remain are shared - each
// warp evaluates them on its own 64-sample chunks of every tile - so that all warps carry the same load (with
// unequal state counts the CTA runs at the pace of its widest warps: the ring of sample tiles couples them).
// MIXED, LEFT == 0: one chunk per warp, the last a.nwide warps own WT+1 states and the others WT, so that H states tile
// any warp count exactly (no idle state slots) and the warp count can be a multiple of the 4 SM sub-partitions.
// vblk / vnblk: index of this CTA among the CTAs of its rank (= blockIdx.x / gridDim.x except in the
// one-GPU emulation of two ranks).
template <int D, int WT, bool MIXED, int LEFT, int TSR>
__device__ __forceinline__ void eval_grad_body(const EvalArgs& a, const int vblk, const int vnblk, unsigned char* smem) {
  constexpr int WTA = (MIXED && LEFT == 0) ? WT + 1 : WT;  // accumulator rows per warp
  constexpr int LA = LEFT > 0 ? LEFT : 1;                  // shared (left-over) states, array extent
  const int H = a.H, S = a.d.S, A = a.d.A;
  const bool roll = a.d.kind == KLERG_DYN_ROLL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const SmemPlan sp = plan_grad<D>(H, S, A, roll, nwarps, WTA + LEFT, TSR);
  float* s_u = (float*)(smem + sp.u);
  float* s_traj = (float*)(smem + sp.traj);
  float* s_dbarr = (float*)(smem + sp.dbarr);
  float* s_P = roll ? (float*)(smem + sp.P) : nullptr;
  float* s_rot = (float*)(smem + sp.rot);
  u64* s_x2 = (u64*)(smem + sp.x2);
  float* s_rx = (float*)(smem + sp.rx);
  float* s_xs = (float*)(smem + sp.xs);
  float* s_tile = (float*)(smem + sp.tile);
  float* s_part = (float*)(smem + sp.part);
  double* s_red = (double*)(smem + sp.red);
  int* s_flag = (int*)(smem + sp.misc);
  unsigned* s_epoch = (unsigned*)(smem + sp.misc) + 1;
  float* s_bsum = (float*)(smem + sp.misc) + 3;
  float* s_ctr = (float*)(s_red + 32 * 2 + 4);  // [D] centre of the trajectory (scaled coordinates)
  constexpr int NF = RowX<D>::NF;
  constexpr int TR = TILE_ROWS(D);  // rows of a sample tile: s_0..s_{D-1}, v -> w, p, |sc|^2
  unsigned* ctrl = ws_fused_ctrl(a.ws);
  void* me = a.peers.mail[a.peers.rank];
  const int world = a.peers.world;

  KLERG_STAMP_DECL;
  KLERG_STAMP(0);
  KLERG_CTA_STAMP(me, vblk, 0);
  // ---- phase 0: rollout, states only (every CTA) ----------------------------------------------------
  float* s_x0 = (float*)(smem + sp.misc) + 4;  // [S] (+ [9] R0)
  // full[b]: the TMA copies of the tile in buffer b have landed
  unsigned long long* s_full = (unsigned long long*)(smem + ((sp.misc + 16 + sizeof(float) * (KLERG_MAX_S + 9) + 16 + 7) & ~(size_t)7));
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < TILE_NB; ++b) mbar_init(&s_full[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!a.independent) pdl_wait_prior_grids();  // the previous launch may have produced this one's inputs
  if (a.gate && __ldcg(a.gate) <= 0) return;
  if (tid == 0) {
    s_epoch[0] = ld_acquire_u32(mb_hdr(me));      // launches so far on this mailbox
    s_epoch[1] = ld_acquire_u32(mb_hdr(me) + 1);  // gather exchanges so far
    // The rollout is the same in every CTA.  The launch's first CTA publishes it; a CTA that starts later (with
    // overlapping launches the first CTA of the next eval starts while this eval's gradient pass is still running
    // everywhere else) copies it instead of redoing it.
    s_flag[0] = vblk != 0 && ld_acquire_u32(mb_hdr(me) + 4 + (s_epoch[0] & 1u)) == s_epoch[0] + 1u;
  }
  for (int e = tid; e < H * A; e += blockDim.x) s_u[e] = a.u[e];
  if (tid < S) s_x0[tid] = a.x0[tid];
  if (a.R0 && tid >= 32 && tid < 41) s_x0[KLERG_MAX_S + tid - 32] = a.R0[tid - 32];
  __syncthreads();
  KLERG_STAMP(8);
  const int n_rot = roll ? rollout_rot_floats(1, H) : 0;
  if (s_flag[0]) {
    const float* ro = mb_ro(me, s_epoch[0] & 1u);
    for (int e = tid; e < (H + 1) * S; e += blockDim.x) s_traj[e] = __ldcg(ro + e);
    for (int e = tid; e < n_rot; e += blockDim.x) s_rot[e] = __ldcg(ro + (H + 1) * S + e);
    __syncthreads();
  } else {
    rollout_block(a.d, a.bar, s_x0, a.R0 ? s_x0 + KLERG_MAX_S : nullptr, s_u, 1, H, s_traj, nullptr, nullptr, s_rot, (float*)s_red,
                  s_bsum, nullptr, 1);
    if (vblk == 0) {
      float* ro = mb_ro(me, s_epoch[0] & 1u);
      for (int e = tid; e < (H + 1) * S; e += blockDim.x) ro[e] = s_traj[e];
      for (int e = tid; e < n_rot; e += blockDim.x) ro[(H + 1) * S + e] = s_rot[e];
      __threadfence();
      __syncthreads();
      if (tid == 0) st_release_u32(mb_hdr(me) + 4 + (s_epoch[0] & 1u), s_epoch[0] + 1u);
    }
  }
  KLERG_STAMP(9);
  const unsigned epoch = s_epoch[0], xc = s_epoch[1];
  if (vblk == 0 && a.traj)
    for (int e = tid; e < (H + 1) * S; e += blockDim.x) a.traj[e] = s_traj[e];
  // pre-step states (Robot.forward, klerg.py:419-431) in both pair forms; the expanded form while the trajectory
  // stays within its radius around the middle state
  StateRows sr;
  sr.x2 = s_x2; sr.rx = s_rx; sr.ctr = s_ctr;
  sr.xform = stage_state_rows<D>(a, s_traj, S, 1, H, 0, s_x2, s_rx, s_xs, s_ctr);
  const bool xform = sr.xform;

  KLERG_STAMP(1);
  KLERG_CTA_STAMP(me, vblk, 1);
  // ---- phase 1: forward pair pass, slice totals -------------------------------------------------
  int64_t lo, hi;
  cta_slice(a.N, a.ld, vblk, vnblk, lo, hi);
  double* s_world = s_red + 32 * 2;  // [2] all ranks
  double* s_in = s_world + 2;        // [2] this CTA
  {
    double tsum = 0.0, tmax = -INFINITY;
    forward_slice<D>(a, sr, H, a.v, lo, hi, tsum, tmax);
    const int kinds[2] = {RED_SUM, RED_MAX};
    double vals[2] = {tsum, tmax};
    block_reduce<2>(kinds, vals, s_red);
    if (tid == 0) {
      s_in[0] = vals[0];
      s_in[1] = vals[1];
    }
  }
  KLERG_STAMP(2);
  KLERG_CTA_STAMP(me, vblk, 2);
  // The first sample tiles of the gradient pass (samples, this CTA's own v, p) do not depend on the grid-wide
  // totals: their TMA copies are issued now and land while the CTAs meet.
  const int ts = a.ts;
  const int nt = (int)((hi - lo + ts - 1) / ts);       // tiles per sweep of this CTA's slice
  const int g_total = a.K * a.rounds * nt;             // the launch's whole tile sequence: (target, round, tile)
  // Tile g of the sequence goes to ring buffer g % NB; one thread issues its D+2 row copies as TMA bulk copies
  // (contiguous runs, no descriptors) that complete on the buffer's `full` barrier.
  auto issue_tile = [&](int g) {
    const int b = g & (TILE_NB - 1), sidx = g / nt, t = g - sidx * nt;
    const float* p_g = a.p + (int64_t)(sidx / a.rounds) * a.p_stride;
    float* buf = s_tile + (size_t)b * TR * TSR;
    const int64_t base = lo + (int64_t)t * ts;
    const unsigned bytes = 4u * (unsigned)min((int64_t)ts, hi - base);  // multiple of 16
    fence_proxy_async();  // the buffer was last written with ordinary stores (importance ratio, padding)
    mbar_expect_tx(&s_full[b], (D + 2) * bytes);
#pragma unroll
    for (int d = 0; d < D; ++d) tma_bulk_g2s(buf + (size_t)d * TSR, a.packed + (int64_t)d * a.ld + base, bytes, &s_full[b]);
    tma_bulk_g2s(buf + (size_t)D * TSR, a.v + base, bytes, &s_full[b]);
    tma_bulk_g2s(buf + (size_t)(D + 1) * TSR, p_g + base, bytes, &s_full[b]);
  };
  fence_proxy_async();  // v was written with ordinary stores and is read back by TMA
  __syncthreads();
  if (tid == 0 && g_total > 0) issue_tile(0);
  // ---- meeting (1): {sum, max} over every CTA of every rank ----------------------------------------
  // Slots of this launch's first gather exchange (number xc) are about to be reused from exchange xc - 2: wait until
  // that one has been consumed (always true in practice; makes the slot reuse safe by construction).
  if (tid == 0) ll_wait_exchange_free(me, xc, ctrl);
  ll_allreduce(a.peers, vblk, vnblk, epoch, 1u, 2, 0x2u, s_in, s_world, (double*)(s_tile + (size_t)TR * TSR), ctrl);
  if (vblk == 0 && tid == 0) {
    // every CTA holds epoch / xc in registers by now: bump the counters for the next launch, which may start
    // as soon as all CTAs have passed this point
    st_release_u32(mb_hdr(me) + 1, xc + (unsigned)a.K);
    st_release_u32(mb_hdr(me), epoch + 1u);
    __threadfence();
  }
  __syncthreads();
  pdl_launch_dependents();
  KLERG_STAMP(3);
  KLERG_CTA_STAMP(me, vblk, 3);
  const double vsum = s_world[0];
  const double vmax = s_world[1];
  if (vblk == 0 && tid == 0 && a.totals) {
    a.totals[0] = vsum;
    a.totals[1] = vmax;
  }
  const float vsum_f = (float)vsum;  // the reference divides by the fp32 sum
  const float maxc_f = (float)fmax(vmax / vsum, (double)a.floor);

  // ---- phase 2: importance ratio + gradient pair pass ----------------------------------------------
  // loop-invariant launch parameters of the pair loop, pinned through shared memory (see pin_params); so are the
  // few scalars only needed again after the pair loop (they would otherwise cost it registers)
  if (tid == 0) {
    s_flag[4 + KLERG_MAX_S + 9] = a.nsub * 64;
    s_flag[4 + KLERG_MAX_S + 10] = vblk;
    s_flag[4 + KLERG_MAX_S + 11] = vnblk;
  }
  __syncthreads();
  const int pb_step = ((volatile int*)s_flag)[4 + KLERG_MAX_S + 9];
  const int HD = H * D, NE = HD + H;  // gather entries: A[t][d] = sum w psi sc_d, then W[t] = sum w psi
  const bool want_kl = a.kl_out != nullptr || a.cost != nullptr;
  double kl_a = 0.0, kl_c = 0.0;

  // Turn tile g (in buffer g & 1) into pair operands, in place, by the whole CTA: v -> importance ratio w = p/q
  // (klerg.py:436); expanded form: samples centred and |sc|^2; rows beyond the slice zeroed.  Two samples per thread
  // and step, as one packed pair.
  auto convert_tile = [&](int g, bool with_kl, int t) {
    const int b = g & (TILE_NB - 1);
    float* buf = s_tile + (size_t)b * TR * TSR;
    const int64_t base = lo + (int64_t)t * ts;
    const int cnt = (int)min((int64_t)ts, hi - base);
    const int cnt64 = (cnt + 63) & ~63;
    float* wrow = buf + (size_t)D * TSR;
    const float* prow = buf + (size_t)(D + 1) * TSR;
    float* nrow = buf + (size_t)(D + 2) * TSR;
    KLERG_SPIN_UNTIL(mbar_try_wait(&s_full[b], (unsigned)(g / TILE_NB) & 1u), ctrl)
    for (int e = 2 * tid; e < cnt64; e += 2 * (int)blockDim.x) {
      const float2 vv = *reinterpret_cast<const float2*>(&wrow[e]);
      const float2 pp = *reinterpret_cast<const float2*>(&prow[e]);
      const bool in0 = e < cnt && base + e < a.N, in1 = e + 1 < cnt && base + e + 1 < a.N;
      const float c0 = fmaxf(__fdividef(vv.x, vsum_f), a.floor), c1 = fmaxf(__fdividef(vv.y, vsum_f), a.floor);
      const float w0 = in0 ? __fdividef(pp.x * maxc_f, c0) : 0.f, w1 = in1 ? __fdividef(pp.y * maxc_f, c1) : 0.f;
      if (with_kl) {
        if (in0) { kl_a += (double)(pp.x * (logf(pp.x) - logf(c0))); kl_c += (double)c0; }
        if (in1) { kl_a += (double)(pp.y * (logf(pp.y) - logf(c1))); kl_c += (double)c1; }
      }
      u64 n2 = pack2(0.f, 0.f);
      if (e >= cnt) {  // cnt is even: the pair lies beyond the slice as a whole
#pragma unroll
        for (int d = 0; d < D; ++d) *reinterpret_cast<u64*>(&buf[(size_t)d * TSR + e]) = pack2(0.f, 0.f);
        if (xform) n2 = pack2(INFINITY, INFINITY);  // w = 0 in the exponent
      } else if (xform) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          u64* sp2 = reinterpret_cast<u64*>(&buf[(size_t)d * TSR + e]);
          const u64 sc = sub2(*sp2, pack2(s_ctr[d], s_ctr[d]));
          *sp2 = sc;
          n2 = fma2(sc, sc, n2);
        }
        // the weight rides in the exponent: |sc|^2 - log2(w), +inf for w = 0 (2^-inf = 0 = w psi)
        n2 = sub2(n2, pack2(__log2f(w0), __log2f(w1)));
      }
      *reinterpret_cast<float2*>(&wrow[e]) = make_float2(w0, w1);
      *reinterpret_cast<u64*>(&nrow[e]) = n2;
    }
  };

  for (int kt = 0; kt < a.K; ++kt) {  // belief targets: the forward pass above is shared, p_k differs
  if (kt > 0 && tid == 0) ll_wait_exchange_free(me, xc + (unsigned)kt, ctrl);  // ordered before the slot stores by the barriers below
  for (int r = 0; r < a.rounds; ++r) {
    int cw = warp % a.nchr, sub = warp / a.nchr;
    int t0 = (r * a.nchr + cw) * WT, my_wt = WT;
    bool active = sub < a.nsub && t0 < H;
    if (MIXED) {
      if (LEFT > 0) {
        t0 = warp * WT;
      } else {
        const int narrow = nwarps - a.nwide;
        const bool wide = warp >= narrow;
        t0 = warp * WT + (wide ? warp - narrow : 0);
        my_wt = WT + (wide ? 1 : 0);
      }
      sub = 0;
      active = true;
    }
    const int left0 = nwarps * WT;  // balanced schedule: first shared state
    u64 acc[WTA][D], wacc[WTA];
#pragma unroll
    for (int k = 0; k < WTA; ++k) {
#pragma unroll
      for (int d = 0; d < D; ++d) acc[k][d] = pack2(0.f, 0.f);
      wacc[k] = pack2(0.f, 0.f);
    }
    // The shared (left-over) states: their accumulators stay in registers across the sweep, their rows are re-read from
    // shared memory for every tile's share of them (the row registers are only live there).
    float* s_left = s_part + (size_t)nwarps * WTA * (D + 1);  // [nwarps][LEFT][D + 1] partials of the shared states
    u64 lacc[LA][D], lwacc[LA];
#pragma unroll
    for (int k = 0; k < LA; ++k) {
#pragma unroll
      for (int d = 0; d < D; ++d) lacc[k][d] = pack2(0.f, 0.f);
      lwacc[k] = pack2(0.f, 0.f);
    }
    const int g_base = (kt * a.rounds + r) * nt;
    // The sweep over the slice's tiles exists once per pair form (the form is uniform over the launch), so that only
    // that form's copy of the warp's states lives in registers: difference form {x', x'} pairs, expanded form
    // -2 xc and |xc|^2.
    auto run_tiles = [&](auto xf_tag) {
      constexpr bool XF = decltype(xf_tag)::value;
      u64 xs2[XF ? 1 : WTA][D];
      float m2x[XF ? WTA : 1][D], x2n[XF ? WTA : 1];
#pragma unroll
      for (int k = 0; k < WTA; ++k) {
        const bool have = active && k < my_wt && t0 + k < H;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          if constexpr (XF)
            m2x[k][d] = have ? s_rx[(t0 + k) * NF + d] : 0.f;
          else
            xs2[k][d] = have ? s_x2[(t0 + k) * Row2<D>::DP + d] : pack2(0.f, 0.f);
        }
        if constexpr (XF) x2n[k] = have ? s_rx[(t0 + k) * NF + D] : 0.f;
      }
      for (int t = 0; t < nt; ++t) {
        const int g = g_base + t, b = g & (TILE_NB - 1);
        convert_tile(g, want_kl && r == 0, t);
        __syncthreads();  // tile g is ready for everyone; everyone is done with tile g - 1
        if (tid == 0 && g + 1 < g_total) issue_tile(g + 1);  // the other buffer: the next tile (or the next sweep's first)
        float* buf = s_tile + (size_t)b * TR * TSR;
        const int64_t base = lo + (int64_t)t * ts;
        const int cnt = (int)min((int64_t)ts, hi - base);
        const int cnt64 = (cnt + 63) & ~63;
        const float* wrow = buf + (size_t)D * TSR;
        const float* nrow = buf + (size_t)(D + 2) * TSR;
        if (active) {
          if constexpr (XF) {
            // warp-uniform: this warp owns WTA or WTA - 1 states
            auto sweep = [&](auto ns_tag) {
              constexpr int NS = decltype(ns_tag)::value;
              for (int pb = sub * 64; pb < cnt64; pb += pb_step) {
                const int i = pb + 2 * lane;
                u64 s2[D];
#pragma unroll
                for (int d = 0; d < D; ++d) s2[d] = *reinterpret_cast<const u64*>(&buf[(size_t)d * TSR + i]);
                const u64 n2 = *reinterpret_cast<const u64*>(&nrow[i]);
                pair_gradient_xw<D, WTA, NS>(m2x, x2n, s2, n2, acc, wacc);
              }
            };
            if (WTA == 1 || my_wt == WTA)
              sweep(std::integral_constant<int, WTA>{});
            else
              sweep(std::integral_constant<int, (WTA > 1 ? WTA - 1 : 1)>{});
          } else {
            for (int pb = sub * 64; pb < cnt64; pb += pb_step) {
              const int i = pb + 2 * lane;
              u64 s2[D];
#pragma unroll
              for (int d = 0; d < D; ++d) s2[d] = *reinterpret_cast<const u64*>(&buf[(size_t)d * TSR + i]);
              const u64 w2 = *reinterpret_cast<const u64*>(&wrow[i]);
              pair_gradient<D, WTA>(xs2, s2, w2, acc, my_wt == WTA);
            }
          }
        }
        if constexpr (LEFT > 0) {
          // the shared states, on this warp's own 64-sample chunks of the tile (chunk c by warp c % nwarps)
          if (left0 < H && warp < (cnt64 >> 6)) {
            u64 lxs2[XF ? 1 : LA][D];
            float lm2x[XF ? LA : 1][D], lx2n[XF ? LA : 1];
#pragma unroll
            for (int k = 0; k < LA; ++k) {
              const bool have = left0 + k < H;
#pragma unroll
              for (int d = 0; d < D; ++d) {
                if constexpr (XF)
                  lm2x[k][d] = have ? s_rx[(left0 + k) * NF + d] : 0.f;
                else
                  lxs2[k][d] = have ? s_x2[(left0 + k) * Row2<D>::DP + d] : pack2(0.f, 0.f);
              }
              if constexpr (XF) lx2n[k] = have ? s_rx[(left0 + k) * NF + D] : 0.f;
            }
            for (int c = warp; c < (cnt64 >> 6); c += nwarps) {
              const int i = (c << 6) + 2 * lane;
              u64 s2[D];
#pragma unroll
              for (int d = 0; d < D; ++d) s2[d] = *reinterpret_cast<const u64*>(&buf[(size_t)d * TSR + i]);
              if constexpr (XF) {
                const u64 n2 = *reinterpret_cast<const u64*>(&nrow[i]);
                pair_gradient_xw<D, LA, LA>(lm2x, lx2n, s2, n2, lacc, lwacc);
              } else {
                const u64 w2 = *reinterpret_cast<const u64*>(&wrow[i]);
                pair_gradient<D, LA>(lxs2, s2, w2, lacc, true);
              }
            }
          }
        }
      }
    };
    if (xform)
      run_tiles(std::true_type{});
    else
      run_tiles(std::false_type{});
    if constexpr (LEFT > 0) {
#pragma unroll
      for (int k = 0; k < LA; ++k) {
#pragma unroll
        for (int d = 0; d <= D; ++d) {
          float x, y;
          unpack2(d < D ? lacc[k][d] : lwacc[k], x, y);
          const float v = warp_sum_f(x + y);
          if (lane == 0) s_left[(warp * LA + k) * (D + 1) + d] = (d < D && !xform) ? -v : v;
        }
      }
    }
    __syncthreads();
    // lanes -> warp sums -> CTA partial for this round's states.  Gather entries are A[t][d] and W[t] with
    // dgdx = gfac (xc W - A); the difference form holds sum w psi (x' - s') directly: A = -acc, W = 0.
#pragma unroll
    for (int k = 0; k < WTA; ++k) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float x, y;
        unpack2(acc[k][d], x, y);
        const float v = warp_sum_f(x + y);
        if (lane == 0) s_part[(warp * WTA + k) * (D + 1) + d] = xform ? v : -v;
      }
      float x, y;
      unpack2(wacc[k], x, y);
      const float v = warp_sum_f(x + y);
      if (lane == 0) s_part[(warp * WTA + k) * (D + 1) + D] = v;
    }
    __syncthreads();
    // this CTA's partial of gather entry e goes to slot [e][vblk] as ONE tagged word (fp32 value + tag)
    {
    const int my_blk = ((volatile int*)s_flag)[4 + KLERG_MAX_S + 10];
    const unsigned xnum_ = s_epoch[1] + (unsigned)kt;
    const int xpar_ = xnum_ & 1u;
    const unsigned xtag_ = ll_tag(xnum_, 0x80u);
    if (MIXED && LEFT > 0) {
      for (int e = tid; e < NE; e += blockDim.x) {
        const int t = e < HD ? e / D : e - HD, d = e < HD ? e - t * D : D;
        float v = 0.f;
        if (t < left0) {
          v = s_part[t * (D + 1) + d];  // (w * WT + k) = t
        } else {
          for (int w = 0; w < nwarps; ++w) v += s_left[(w * LA + (t - left0)) * (D + 1) + d];
        }
        ll_store_f32(mb_gp(me, xpar_, e) + my_blk, v, xtag_);
      }
    } else if (MIXED) {
      const int narrow = nwarps - a.nwide, narrow_states = narrow * WT;
      for (int e = tid; e < NE; e += blockDim.x) {
        const int t = e < HD ? e / D : e - HD, d = e < HD ? e - t * D : D;
        int w, k;
        if (t < narrow_states) {
          w = t / WT;
          k = t - w * WT;
        } else {
          const int tt = t - narrow_states;
          w = narrow + tt / (WT + 1);
          k = tt - (w - narrow) * (WT + 1);
        }
        ll_store_f32(mb_gp(me, xpar_, e) + my_blk, s_part[(w * WTA + k) * (D + 1) + d], xtag_);
      }
    } else
    for (int e = tid; e < a.nchr * WT * (D + 1); e += blockDim.x) {
      const int c = e / (WT * (D + 1)), kd = e - c * (WT * (D + 1));
      const int t = (r * a.nchr + c) * WT + kd / (D + 1), d = kd % (D + 1);
      if (t < H) {
        float v = 0.f;
        for (int sb = 0; sb < a.nsub; ++sb) v += s_part[((sb * a.nchr + c) * WT) * (D + 1) + kd];
        ll_store_f32(mb_gp(me, xpar_, d < D ? t * D + d : HD + t) + my_blk, v, xtag_);
      }
    }
    }
  }
  // (re)read what the pair loop did not keep in registers
  const int vblk_ = ((volatile int*)s_flag)[4 + KLERG_MAX_S + 10], vnblk_ = ((volatile int*)s_flag)[4 + KLERG_MAX_S + 11];
  const unsigned xnum = s_epoch[1] + (unsigned)kt;  // number of this target's gather exchange
  const int xpar = xnum & 1u;
  const unsigned xtag = ll_tag(xnum, 0x80u);
  if (want_kl) {
    const int kinds[2] = {RED_SUM, RED_SUM};
    double vals[2] = {kl_a, kl_c};
    kl_a = kl_c = 0.0;
    block_reduce<2>(kinds, vals, s_red);
    if (tid == 0) {
      ll_store(mb_kl(me, xpar, vblk_), vals[0], xtag);
      ll_store(mb_kl(me, xpar, vblk_) + 2, vals[1], xtag);
    }
  }

  // ---- phase 3: warps spread over the grid add the CTA partials of one gather entry each (fixed order) and
  //      hand the sum to every rank; the finisher CTA collects the sums of all ranks and runs the adjoint ----
  KLERG_STAMP(4);
  KLERG_CTA_STAMP(me, vblk_, 4);
  for (int e = vblk_ + vnblk_ * warp; e < NE; e += vnblk_ * nwarps) {
    const u64* row = mb_gp(me, xpar, e);
    double v = 0.0;
    for (int b = lane; b < vnblk_; b += 32) {
      float x = 0.f;
      KLERG_SPIN_UNTIL(ll_try_load_f32(row + b, xtag, x), ctrl)
      v += (double)x;
    }
    v = warp_reduce(RED_SUM, v);
    if (lane < world) ll_store(mb_gb(a.peers.mail[lane], xpar, a.peers.rank) + 2 * e, v, xtag);
  }
  KLERG_CTA_STAMP(me, vblk_, 5);
  if (vblk_ != vnblk_ - 1) continue;
  KLERG_STAMP(5);
  if (kt == 0) {
    // what only the adjoint needs: linearisation blocks, dbarr, barrier sum
    __syncthreads();
    rollout_block(a.d, a.bar, s_x0, a.R0 ? s_x0 + KLERG_MAX_S : nullptr, s_u, 1, H, s_traj, s_dbarr, s_P, s_rot, (float*)s_red,
                  s_bsum, nullptr, 2);
  }
  __syncthreads();
  // the finisher's own scratch (the tile ring may already hold tiles of the next target)
  double* s_val = (double*)(smem + sp.adj);      // [NE + 2]
  float* s_g = (float*)(s_val + NE + 2);         // [H][S]
  float* s_scr = s_g + H * S;                    // adjoint scratch
  if (want_kl) {
    // KL terms of this rank: CTA partials in CTA order, then to every rank like a gather entry
    double* s_stage = (double*)(((uintptr_t)(s_scr + adjoint_scratch_floats(H, A)) + 7) & ~(uintptr_t)7);  // [2 vnblk] + [2] + [32]
    for (int idx = tid; idx < vnblk_ * 2; idx += blockDim.x) {
      double x = 0.0;
      KLERG_SPIN_UNTIL(ll_try_load(mb_kl(me, xpar, idx >> 1) + 2 * (idx & 1), xtag, x), ctrl)
      s_stage[idx] = x;
    }
    __syncthreads();
    ll_reduce_staged(s_stage, vnblk_, 2, 0u, s_stage + 2 * LL_MAXBLK, s_stage + 2 * LL_MAXBLK + 2);
    if (tid < world * 2) {
      const int rr = tid >> 1, i = tid & 1;
      ll_store(mb_gb(a.peers.mail[rr], xpar, a.peers.rank) + 2 * (NE + i), s_stage[2 * LL_MAXBLK + i], xtag);
    }
  }
  for (int e = tid; e < NE + (want_kl ? 2 : 0); e += blockDim.x) {
    double v = 0.0;
    for (int rr = 0; rr < world; ++rr) {
      double x = 0.0;
      KLERG_SPIN_UNTIL(ll_try_load(mb_gb(me, xpar, rr) + 2 * e, xtag, x), ctrl)
      v += x;
    }
    s_val[e] = v;
  }
  if (!want_kl && tid == 0) {
    s_val[NE] = 0.0;
    s_val[NE + 1] = 1.0;
  }
  KLERG_STAMP(6);
  for (int e = tid; e < H * S; e += blockDim.x) s_g[e] = 0.f;
  __syncthreads();
  for (int e = tid; e < HD; e += blockDim.x) {
    const int t = e / D, d = e - t * D;
    // dgdx_t[d] = gfac_d * sum_i w psi (x'_td - s'_id) = gfac_d * (xc_td W_t - A_td), formed in double
    s_g[t * S + a.k.explr[d]] = (float)(((double)s_xs[e] * s_val[HD + t] - s_val[e]) * (double)a.k.gfac[d]);
  }
  __syncthreads();
  for (int e = tid; e < H * S; e += blockDim.x) {
    const float g = s_g[e];
    a.dgdx[(size_t)kt * H * S + e] = g;
    s_g[e] = g - s_dbarr[e];
  }
  __syncthreads();
  adjoint_block(a.d, a.ap, H, s_g, s_P, s_traj, s_u, s_scr, a.du + (size_t)kt * H * A, a.djdlam + (size_t)kt * H,
                a.u_star + (size_t)kt * H * A);
  if (tid == 0) {
    const double sa = s_val[NE], sc = s_val[NE + 1];
    if (a.kl_out) {
      a.kl_out[2 * kt] = sa;
      a.kl_out[2 * kt + 1] = sc;
    }
    if (a.cost) {
      const double spv = a.p_stats[kt];
      // KL of the PRE-step footprint (what backward() differentiates) + barrier of the post-step states
      a.cost[kt] = (float)(sa / spv - log(spv) + log(sc)) + *s_bsum;
    }
    // this exchange is consumed: its slots may be reused two exchanges from now (in order, one finisher at a time)
    unsigned* xdone = mb_hdr(me) + 2;
    KLERG_SPIN_UNTIL(ld_acquire_u32(xdone) == xnum, ctrl)
    __threadfence();
    st_release_u32(xdone, xnum + 1u);
    if (kt == a.K - 1 && a.fault_out) *a.fault_out = ctrl[5] ? 1.f : 0.f;
#ifdef KLERG_STAMPS
    // phase stamps of the finisher (SM cycles since its start): profiling aid
    KLERG_STAMP(7);
    KLERG_CTA_STAMP(me, vblk_, 6);
    long long* dbg = (long long*)(ctrl + 16);
    for (int i = 0; i < 10; ++i) dbg[i] = stamp[i] - stamp[0];
    for (int i = 0; i < 8; ++i) dbg[10 + i] = g_ro_stamp[i] - g_ro_stamp[0];
#endif
  }
  __syncthreads();
  }  // targets
}

template <int D, int WT, int MAXT, bool MIXED, int LEFT>
__global__ void __launch_bounds__(MAXT) eval_grad_kernel(const __grid_constant__ EvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  eval_grad_body<D, WT, MIXED, LEFT, TS_ROW>(a, (int)blockIdx.x, (int)gridDim.x, smem);
}
// two ranks on one GPU (tests): CTAs [0, nb) act as rank 0 with a0, CTAs [nb, 2 nb) as rank 1 with a1
template <int D, int WT, int MAXT, bool MIXED, int LEFT>
__global__ void __launch_bounds__(MAXT) eval_grad_emu_kernel(const __grid_constant__ EvalArgs a0,
                                                             const __grid_constant__ EvalArgs a1, const int nb) {
  extern __shared__ __align__(16) unsigned char smem[];
  if ((int)blockIdx.x < nb)
    eval_grad_body<D, WT, MIXED, LEFT, TS_ROW>(a0, (int)blockIdx.x, nb, smem);
  else
    eval_grad_body<D, WT, MIXED, LEFT, TS_ROW>(a1, (int)blockIdx.x - nb, nb, smem);
}

// KL partials of G candidates over the slice [lo, hi): sa[g] = sum_i p_i (log p_i - log c_i), sc[g] = sum_i c_i
// (klerg.py:694-699 in closed form), from v[g][i] = q_base + q_iter and the candidates' world totals {sum, max}.
// Four samples per thread-iteration (128-bit loads of v and p), reciprocal of the normaliser and lg2-based
// logarithms: the pass is instruction-bound (IEEE division + logf cost ~60 instructions per sample and candidate,
// this form ~12); the cost changes by < 1e-6 relative, far inside the 1e-4 parity tolerance.
__device__ __forceinline__ void kl_pass(const EvalArgs& a, int G, const float* vbase, const double* s_world, int64_t lo, int64_t hi,
                                        double (&sa)[FUSED_MAXG], double (&sc)[FUSED_MAXG]) {
  const int tid = threadIdx.x;
  float rvs[FUSED_MAXG], maxc[FUSED_MAXG];
#pragma unroll
  for (int g = 0; g < FUSED_MAXG; ++g) {
    sa[g] = sc[g] = 0.0;
    rvs[g] = maxc[g] = 1.f;
    if (g < G) {
      const double vsum = s_world[2 * g], vmax = s_world[2 * g + 1];
      const float vs = (float)vsum;
      rvs[g] = 1.f / vs;
      maxc[g] = fmaxf((float)vmax / vs, a.floor);
    }
  }
  const int64_t hiN = hi < a.N ? hi : a.N;
  for (int64_t i0 = lo + (int64_t)tid * 4; i0 < hiN; i0 += (int64_t)blockDim.x * 4) {
    float pv[4], lp[4];
    if (i0 + 3 < a.N) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(a.p + i0));
      pv[0] = t.x; pv[1] = t.y; pv[2] = t.z; pv[3] = t.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) pv[q] = (i0 + q < a.N) ? a.p[i0 + q] : 1.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (pv[q] != pv[q]) pv[q] = 1e-6f;
      lp[q] = __logf(pv[q]);
    }
#pragma unroll
    for (int g = 0; g < FUSED_MAXG; ++g) {
      if (g < G) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(vbase + (size_t)g * a.ld + i0));
        const float vv[4] = {t.x, t.y, t.z, t.w};
        float fa = 0.f, fc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (i0 + q < hiN) {
            float c = fmaxf(vv[q] * rvs[g], a.floor);
            if (c != c) c = 1e-6f * maxc[g];  // cost_norm: NaN in q -> 1e-6 (q = c / max c)
            fa = fmaf(pv[q], lp[q] - __logf(c), fa);
            fc += c;
          }
        }
        sa[g] += (double)fa;
        sc[g] += (double)fc;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// cost eval of G <= FUSED_MAXG candidates
// ---------------------------------------------------------------------------
template <int D>
__host__ __device__ inline SmemPlan plan_cost(int G, int H, int S, int A, bool roll) {
  SmemPlan p{};
  size_t o = 0;
  p.u = o;    o = align16(o + sizeof(float) * (size_t)G * H * A);
  p.traj = o; o = align16(o + sizeof(float) * (size_t)G * (H + 1) * S);
  p.x2 = o;   o = align16(o + sizeof(u64) * (size_t)G * H * Row2<D>::DP);
  p.rx = o;   o = align16(o + sizeof(float) * (size_t)G * H * RowX<D>::NF);
  p.tile = o; o = align16(o + (roll ? sizeof(float) * rollout_rot_floats(G, H) : 0));
  p.red = o;  o = align16(o + sizeof(double) * (32 * 2 * FUSED_MAXG + 4 * FUSED_MAXG));
  p.misc = o; o = align16(o + 64 + sizeof(float) * FUSED_MAXG);
  p.ll = o;   o = align16(o + LL_SCRATCH_BYTES);
  p.total = o;
  return p;
}

template <int D>
__device__ __forceinline__ void eval_cost_body(const EvalArgs& a, const int vblk, const int vnblk, unsigned char* smem) {
  const int H = a.H, S = a.d.S, A = a.d.A;
  const int tid = threadIdx.x;
  int G = a.G;
  if (a.gate) {
    pdl_wait_prior_grids();
    const int g = __ldcg(a.gate);
    if (g <= 0) return;
    if (g < G) G = g;
  }
  const SmemPlan sp = plan_cost<D>(G, H, S, A, a.d.kind == KLERG_DYN_ROLL);
  float* s_u = (float*)(smem + sp.u);
  float* s_traj = (float*)(smem + sp.traj);
  u64* s_x2 = (u64*)(smem + sp.x2);
  float* s_rx = (float*)(smem + sp.rx);
  double* s_red = (double*)(smem + sp.red);
  double* s_world = s_red + 32 * 2 * FUSED_MAXG;  // [2G] all ranks
  double* s_in = s_world + 2 * FUSED_MAXG;        // [2G] this CTA
  unsigned* s_epoch = (unsigned*)(smem + sp.misc) + 1;
  float* s_ctr = (float*)(smem + sp.misc) + 4;    // [D]
  float* s_bsum = (float*)(smem + sp.misc + 64);
  double* s_ll = (double*)(smem + sp.ll);
  unsigned* ctrl = ws_fused_ctrl(a.ws);
  void* me = a.peers.mail[a.peers.rank];
  const int world = a.peers.world;

  if (!a.independent) pdl_wait_prior_grids();
  if (tid == 0) {
    s_epoch[0] = ld_acquire_u32(mb_hdr(me));
    s_epoch[1] = ld_acquire_u32(mb_hdr(me) + 1);
  }
  for (int e = tid; e < G * H * A; e += blockDim.x) s_u[e] = a.u[e];
  __syncthreads();
  rollout_block(a.d, a.bar, a.x0, a.R0, s_u, G, H, s_traj, nullptr, nullptr, (float*)(smem + sp.tile), (float*)s_red,
                s_bsum, nullptr);
  const unsigned epoch = s_epoch[0], xc = s_epoch[1];
  if (vblk == 0 && a.traj)
    for (int e = tid; e < G * (H + 1) * S; e += blockDim.x) a.traj[e] = s_traj[e];
  // post-step states of every candidate (klerg.py:688-691) in both pair forms, centred on candidate 0's middle state
  StateRows sr;
  sr.x2 = s_x2; sr.rx = s_rx; sr.ctr = s_ctr;
  sr.xform = stage_state_rows<D>(a, s_traj + S, S, G, H, (size_t)(H + 1) * S, s_x2, s_rx, nullptr, s_ctr);

  int64_t lo, hi;
  cta_slice(a.N, a.ld, vblk, vnblk, lo, hi);
  if (narrow_pairs(lo, hi))
    forward_candidates<D, 1>(a, sr, G, H, lo, hi, s_red, s_in, a.v);
  else
    forward_candidates<D, 2>(a, sr, G, H, lo, hi, s_red, s_in, a.v);
  // meeting (1): {sum, max} per candidate over every CTA of every rank
  if (tid == 0) ll_wait_exchange_free(me, xc, ctrl);
  ll_allreduce(a.peers, vblk, vnblk, epoch, 1u, 2 * G, 0xAAAAu, s_in, s_world, s_ll, ctrl);
  if (vblk == 0 && tid == 0) {
    st_release_u32(mb_hdr(me) + 1, xc + 1u);
    st_release_u32(mb_hdr(me), epoch + 1u);
    __threadfence();
  }
  __syncthreads();
  pdl_launch_dependents();
  if (vblk == 0 && tid < 2 * G && a.totals) a.totals[tid] = s_world[tid];

  double sa[FUSED_MAXG], sc[FUSED_MAXG];
  kl_pass(a, G, a.v, s_world, lo, hi, sa, sc);
  block_reduce_pairs(G, RED_SUM, sa, sc, s_red);
  // gather: the CTA's 2G KL terms as tagged values; the finisher adds them in CTA order, then in rank order
  const int xpar = xc & 1u;
  const unsigned xtag = ll_tag(xc, 0x80u);
  if (tid == 0) {
#pragma unroll
    for (int g = 0; g < FUSED_MAXG; ++g)
      if (g < G) {
        ll_store(mb_kl(me, xpar, vblk) + 4 * g, sa[g], xtag);
        ll_store(mb_kl(me, xpar, vblk) + 4 * g + 2, sc[g], xtag);
      }
  }
  if (vblk != vnblk - 1) return;
  const int nv = 2 * G;
  for (int idx = tid; idx < vnblk * nv; idx += blockDim.x) {
    const int b = idx / nv, i = idx - b * nv;
    double x = 0.0;
    KLERG_SPIN_UNTIL(ll_try_load(mb_kl(me, xpar, b) + 2 * i, xtag, x), ctrl)
    s_ll[idx] = x;
  }
  __syncthreads();
  double* s_val = s_world;  // [2G] (the totals are in registers by now)
  ll_reduce_staged(s_ll, vnblk, nv, 0u, s_val, s_ll + LL_BUF_VALS);
  if (world > 1) {
    for (int t = tid; t < world * nv; t += blockDim.x) {
      const int r = t / nv, i = t - r * nv;
      ll_store(mb_gb(a.peers.mail[r], xpar, a.peers.rank) + 2 * i, s_val[i], xtag);
    }
    __syncthreads();
    if (tid < nv) {
      double v = 0.0;
      for (int r = 0; r < world; ++r) {
        double x = 0.0;
        KLERG_SPIN_UNTIL(ll_try_load(mb_gb(me, xpar, r) + 2 * tid, xtag, x), ctrl)
        v += x;
      }
      s_val[tid] = v;
    }
    __syncthreads();
  }
  if (tid < G) {
    const double spv = a.p_stats[0];
    const double dkl = s_val[2 * tid] / spv - log(spv) + log(s_val[2 * tid + 1]);
    a.cost[tid] = (float)dkl + s_bsum[tid];
  }
  if (tid == 0) {
    unsigned* xdone = mb_hdr(me) + 2;
    KLERG_SPIN_UNTIL(ld_acquire_u32(xdone) == xc, ctrl)
    __threadfence();
    st_release_u32(xdone, xc + 1u);
    if (a.fault_out) *a.fault_out = ctrl[5] ? 1.f : 0.f;
  }
}

template <int D>
__global__ void __launch_bounds__(512) eval_cost_kernel(const __grid_constant__ EvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  eval_cost_body<D>(a, (int)blockIdx.x, (int)gridDim.x, smem);
}
template <int D>
__global__ void __launch_bounds__(512) eval_cost_emu_kernel(const __grid_constant__ EvalArgs a0, const __grid_constant__ EvalArgs a1,
                                                            const int nb) {
  extern __shared__ __align__(16) unsigned char smem[];
  if ((int)blockIdx.x < nb)
    eval_cost_body<D>(a0, (int)blockIdx.x, nb, smem);
  else
    eval_cost_body<D>(a1, (int)blockIdx.x - nb, nb, smem);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct GradSchedule {
  int wt, nwarps, nchr, nsub, rounds;
  double eff;
  bool mixed;
  int nwide;
  int left;  // balanced mixed schedule: states shared by all warps (H - nwarps * wt)
};

static int grad_max_warps(int D) { return D <= 3 ? 20 : (D == 4 ? 16 : 17); }

// Choose states-per-warp WT and the warp grid so that (states x sample sub-streams) tiles the
// CTA's warps with as few idle slots as possible.
static GradSchedule plan_schedule(int D, int H) {
  if (D >= 5 && (g_fused_opt.mixed_warps == 12 || g_fused_opt.mixed_warps == 0)) {
    // balanced schedule on 12 warps (3 per SM sub-partition, up to 168 registers): every warp owns 4 states, the
    // H - 48 <= 2 states that remain are shared (each warp takes them on its own sample chunks of every tile, with
    // their rows re-read per tile, their accumulators kept in registers).  Equal state counts = no warp waits for
    // a wider one at the tile barrier (the 16-warp split of H = 50 into 14 x 3 + 2 x 4 loses ~6 % there): 616 vs
    // 647 us per c4 eval.  KLERG_OPT_MIXED_WARPS = 1 selects the 16-warp split instead.
    const int nw = 12, q = H / nw, r = H % nw;
    if (q == 4 && r <= 2) {
      GradSchedule m{};
      m.wt = q; m.nwarps = nw; m.nchr = nw; m.nsub = 1; m.rounds = 1; m.eff = 1.0; m.mixed = true; m.nwide = 0; m.left = r;
      return m;
    }
  }
  if (D >= 5 && g_fused_opt.mixed_warps == 16) {
    // balanced schedule on 16 warps: every warp owns q states, the r <= 2 states that remain are shared (each warp
    // takes them on its own 1/16 of a tile's samples) - no warp with one state more that the others wait for at
    // every tile barrier
    const int nw = 16, q = H / nw, r = H % nw;
    if (q == 3 && r >= 1 && r <= 2) {
      GradSchedule m{};
      m.wt = q; m.nwarps = nw; m.nchr = nw; m.nsub = 1; m.rounds = 1; m.eff = 1.0; m.mixed = true; m.nwide = 0; m.left = r;
      return m;
    }
  }
  if (D >= 4) {
    // mixed schedule: 16 warps (4 per SM sub-partition, 128 registers), H = q*16 + r -> r warps own q+1 states.
    // Needs q >= 2 (fewer states per warp would re-read the staged samples too often for the shared-memory bandwidth).
    const int nw = 16, q = H / nw, r = H % nw, wtmax = D == 4 ? 5 : 4;
    if (q >= 2 && r > 0 && q + 1 <= wtmax) {
      GradSchedule m{};
      m.wt = q; m.nwarps = nw; m.nchr = nw; m.nsub = 1; m.rounds = 1; m.eff = 1.0; m.mixed = true; m.nwide = r;
      return m;
    }
    if (r == 0 && q >= 3 && q <= wtmax) {
      // H a multiple of 16: every warp is a "wide" warp of the (q - 1)-state kernel (the same 512-thread, 128-register
      // kernel; the generic schedule below would take the 17-warp kernel with 96 registers)
      GradSchedule m{};
      m.wt = q - 1; m.nwarps = nw; m.nchr = nw; m.nsub = 1; m.rounds = 1; m.eff = 1.0; m.mixed = true; m.nwide = nw;
      return m;
    }
  }
  const int maxw = grad_max_warps(D);
  const int wts_small[5] = {5, 4, 3, 2, 1};
  const int wts_big[3] = {3, 2, 1};
  const int* wts = D <= 4 ? wts_small : wts_big;
  const int nw = D <= 4 ? 5 : 3;
  GradSchedule best{};
  best.eff = -1.0;
  for (int i = 0; i < nw; ++i) {
    const int wt = wts[i];
    const int nch = (H + wt - 1) / wt;
    GradSchedule s{};
    s.wt = wt;
    if (nch <= maxw) {
      s.rounds = 1;
      s.nchr = nch;
    } else {
      s.rounds = (nch + maxw - 1) / maxw;
      s.nchr = (nch + s.rounds - 1) / s.rounds;
    }
    s.nsub = maxw / s.nchr;
    s.nwarps = s.nchr * s.nsub;
    // useful pair slots / issued pair slots, discounted when few warps are resident or the tile is restaged
    s.eff = (double)H / ((double)s.rounds * s.nchr * wt) * (0.5 + 0.5 * s.nwarps / maxw) / (1.0 + 0.02 * (s.rounds - 1));
    if (s.eff > best.eff + 1e-9) best = s;
  }
  return best;
}


template <int DUMMY>
__global__ void fused_probe_kernel() {}

// Launch of a fused eval: programmatic stream serialization (the next fused eval may start while this one's
// finisher is still busy) and, where the driver accepts the combination, the cooperative attribute (all CTAs
// co-resident: they wait for one another).  Without it co-residency holds by construction: the grid never
// exceeds what the device holds at once (pick_grid / resident_ctas) and a launch only waits for CTAs of earlier
// launches, which never wait for it.
template <typename K, typename... Args>
static int fused_launch(K kernel, int nblk, int nthreads, size_t smem, cudaStream_t stream, const char* what, bool force_coop,
                        Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)nblk);
  cfg.blockDim = dim3((unsigned)nthreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  const bool pdl = g_fused_opt.pdl && !force_coop;
  if (pdl && g_fused_opt.coop_probe < 0) {
    // once per process: may the two attributes be combined?
    cudaLaunchConfig_t pc{};
    pc.gridDim = dim3(1);
    pc.blockDim = dim3(32);
    pc.stream = stream;
    cudaLaunchAttribute pa[2];
    pa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pa[0].val.programmaticStreamSerializationAllowed = 1;
    pa[1].id = cudaLaunchAttributeCooperative;
    pa[1].val.cooperative = 1;
    pc.attrs = pa;
    pc.numAttrs = 2;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    if (cs == cudaStreamCaptureStatusNone) {
      const cudaError_t pe = cudaLaunchKernelEx(&pc, fused_probe_kernel<0>);
      g_fused_opt.coop_probe = pe == cudaSuccess ? 1 : 0;
      cudaGetLastError();
    }
  }
  if (pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (!pdl || g_fused_opt.coop_probe == 1) {
    attrs[na].id = cudaLaunchAttributeCooperative;
    attrs[na].val.cooperative = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) {
    set_error("%s: launch failed (grid %d x %d threads, %zu B smem): %s", what, nblk, nthreads, smem, cudaGetErrorString(e));
    cudaGetLastError();
    return -4;
  }
  return check_launch(what);
}

// resident CTAs per SM for a kernel/block/smem combination (cached per kernel pointer + shape).
// The dynamic shared-memory limit of a kernel is only ever raised.
template <typename K>
static int resident_ctas(K kernel, int nthreads, size_t smem) {
  struct Key {
    const void* k;
    int t;
    size_t s;
    int n;
  };
  static Key cache[64];
  static int ncache = 0;
  static const void* raised_k[64];
  static size_t raised_s[64];
  static int nraised = 0;
  for (int i = 0; i < ncache; ++i)
    if (cache[i].k == (const void*)kernel && cache[i].t == nthreads && cache[i].s == smem) return cache[i].n;
  if (smem > 48 * 1024) {
    int slot = -1;
    for (int i = 0; i < nraised; ++i)
      if (raised_k[i] == (const void*)kernel) slot = i;
    if (slot < 0 && nraised < 64) {
      slot = nraised++;
      raised_k[slot] = (const void*)kernel;
      raised_s[slot] = 0;
    }
    if (slot < 0 || raised_s[slot] < smem) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (slot >= 0) raised_s[slot] = smem;
    }
  }
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, nthreads, smem) != cudaSuccess) n = 0;
  cudaGetLastError();
  if (ncache < 64) cache[ncache++] = Key{(const void*)kernel, nthreads, smem, n};
  return n;
}


// CTAs of a fused eval: at least min_samples_per_cta samples each, at most what the device holds at once.  A grid
// that would fill every SM leaves one free: the finisher CTA of the previous eval is still running when the next
// eval's CTAs start (programmatic dependent launch), and with a free SM none of them has to wait for it.
static int pick_grid(int64_t N, int per_sm, int min_samples_per_cta) {
  int64_t nblk = (N + min_samples_per_cta - 1) / min_samples_per_cta;
  int64_t cap = (int64_t)sm_count() * per_sm;
  if (g_fused_opt.pdl && cap > 8) cap -= 1;
  if (cap > LL_MAXBLK) cap = LL_MAXBLK;
  if (g_fused_opt.grid_limit > 0 && cap > g_fused_opt.grid_limit) cap = g_fused_opt.grid_limit;
  if (nblk > cap) nblk = cap;
  if (nblk < 1) nblk = 1;
  return (int)nblk;
}

// samples that decide the grid: the largest shard, so that every rank launches the same number of CTAs
static int64_t grid_samples(const EvalArgs& a, int64_t n_max) { return n_max > a.N ? n_max : a.N; }

template <int D, int WT, int MW, int LEFT>
static int launch_grad_emu(const EvalArgs& a0, const EvalArgs& a1, int nblk, int nthreads, size_t smem, cudaStream_t stream) {
  constexpr bool MIXED = MW > 0;
  constexpr int MAXT = MIXED ? MW * 32 : (D <= 3 ? 20 : (D == 4 ? 16 : 17)) * 32;
  auto kernel = eval_grad_emu_kernel<D, WT, MAXT, MIXED, LEFT>;
  if (resident_ctas(kernel, nthreads, smem) < 1) { set_error("emulated eval_gradient: kernel does not fit on an SM"); return -4; }
  return fused_launch(kernel, 2 * nblk, nthreads, smem, stream, "eval_grad_emu_kernel", true, a0, a1, nblk);
}

template <int D, int WT, int MW = 0, int LEFT = 0>
static int launch_grad_wt(EvalArgs& a, const GradSchedule& s, int64_t n_max, cudaStream_t stream) {
  constexpr bool MIXED = MW > 0;
  constexpr int MAXT = MIXED ? MW * 32 : (D <= 3 ? 20 : (D == 4 ? 16 : 17)) * 32;
  auto kernel = eval_grad_kernel<D, WT, MAXT, MIXED, LEFT>;
  const int nthreads = s.nwarps * 32;
  const bool roll = a.d.kind == KLERG_DYN_ROLL;
  // tile: up to TS_ROW samples, but no more than one CTA's slice at full grid
  int64_t per = (a.N + sm_count() - 1) / sm_count();
  int ts = TS_ROW;
  while (ts > 64 && ts / 2 >= per) ts /= 2;
  a.ts = ts;
  a.nchr = s.nchr; a.nsub = s.nsub; a.rounds = s.rounds; a.nwide = s.nwide;
  const SmemPlan sp = plan_grad<D>(a.H, a.d.S, a.d.A, roll, s.nwarps, ((MIXED && LEFT == 0) ? WT + 1 : WT) + LEFT, TS_ROW);
  if (sp.total > 220 * 1024) { set_error("eval_gradient: horizon too long for shared-memory staging"); return -1; }
  const int per_sm = resident_ctas(kernel, nthreads, sp.total);
  if (per_sm < 1) { set_error("eval_gradient: kernel does not fit on an SM (threads=%d smem=%zu)", nthreads, sp.total); return -4; }
  const int nblk = pick_grid(grid_samples(a, n_max), 1, 128);
  if (g_emu.active) {
    const int r = a.peers.rank;
    if (a.peers.world != 2 || r < 0 || r > 1) { set_error("emulation records world-2 launches only"); return -1; }
    g_emu.args[r] = a;
    g_emu.have[r] = 1;
    g_emu.launch = launch_grad_emu<D, WT, MW, LEFT>;
    g_emu.nblk = nblk; g_emu.nthreads = nthreads; g_emu.smem = sp.total;
    return 0;
  }
  return fused_launch(kernel, nblk, nthreads, sp.total, stream, "eval_grad_kernel", false, a);
}

template <int D>
int launch_grad_d(EvalArgs& a, int64_t n_max, cudaStream_t stream) {
  const GradSchedule s = plan_schedule(D, a.H);
  if constexpr (D >= 4) {
    if (s.mixed) {
      if constexpr (D >= 5) {
        if (s.nwarps == 12 && s.wt == 4) return launch_grad_wt<D, 4, 12, 2>(a, s, n_max, stream);
        if (s.nwarps == 16 && s.wt == 3 && s.left > 0) return launch_grad_wt<D, 3, 16, 2>(a, s, n_max, stream);
      }
      if (s.wt == 2) return launch_grad_wt<D, 2, 16>(a, s, n_max, stream);
      if (s.wt == 3) return launch_grad_wt<D, 3, 16>(a, s, n_max, stream);
      if constexpr (D == 4) {
        if (s.wt == 4) return launch_grad_wt<D, 4, 16>(a, s, n_max, stream);
      }
      set_error("eval_gradient: no mixed schedule for D=%d H=%d", D, a.H);
      return -2;
    }
  }
  switch (s.wt) {
    case 1: return launch_grad_wt<D, 1>(a, s, n_max, stream);
    case 2: return launch_grad_wt<D, 2>(a, s, n_max, stream);
    case 3: return launch_grad_wt<D, 3>(a, s, n_max, stream);
    case 4:
      if constexpr (D <= 4) return launch_grad_wt<D, 4>(a, s, n_max, stream);
      break;
    case 5:
      if constexpr (D <= 4) return launch_grad_wt<D, 5>(a, s, n_max, stream);
  }
  set_error("eval_gradient: no schedule for D=%d H=%d", D, a.H);
  return -2;
}

template <int D>
static int launch_cost_emu(const EvalArgs& a0, const EvalArgs& a1, int nblk, int nthreads, size_t smem, cudaStream_t stream) {
  auto kernel = eval_cost_emu_kernel<D>;
  if (resident_ctas(kernel, nthreads, smem) < 1) { set_error("emulated eval_costs: kernel does not fit on an SM"); return -4; }
  return fused_launch(kernel, 2 * nblk, nthreads, smem, stream, "eval_cost_emu_kernel", true, a0, a1, nblk);
}

template <int D>
int launch_cost_d(EvalArgs& a, int64_t n_max, cudaStream_t stream) {
  auto kernel = eval_cost_kernel<D>;
  const SmemPlan sp = plan_cost<D>(a.G, a.H, a.d.S, a.d.A, a.d.kind == KLERG_DYN_ROLL);
  if (sp.total > 200 * 1024) { set_error("eval_costs: G*H too large for shared-memory staging"); return -1; }
  int nthreads = 512;
  const int per_sm = resident_ctas(kernel, nthreads, sp.total);
  if (per_sm < 1) { set_error("eval_costs: kernel does not fit on an SM"); return -4; }
  const int nblk = pick_grid(grid_samples(a, n_max), 1, 256);
  if (g_emu.active) {
    const int r = a.peers.rank;
    if (a.peers.world != 2 || r < 0 || r > 1) { set_error("emulation records world-2 launches only"); return -1; }
    g_emu.args[r] = a;
    g_emu.have[r] = 1;
    g_emu.launch = launch_cost_emu<D>;
    g_emu.nblk = nblk; g_emu.nthreads = nthreads; g_emu.smem = sp.total;
    return 0;
  }
  return fused_launch(kernel, nblk, nthreads, sp.total, stream, "eval_cost_kernel", false, a);
}

}  // namespace klerg

// Fused KL-ergodic evals: ONE cooperative launch per eval (sm_100a).
//
//   eval_grad_kernel : Robot.forward + footprint + renormalize + importance ratio +
//                      kldiv_grad_vec for all H steps + Robot.backward   (klerg.py:409-450, 505-523)
//                      -> du, djdlam, u*, dgdx, KL cost of the plan
//   eval_cost_kernel : Robot.get_cost for G <= 8 candidate control sequences (klerg.py:686-710;
//                      the <= 5 line-search windows of klerg.py:712-751 are one launch)
//
// Every CTA redoes the tiny rollout in shared memory (no broadcast needed), owns a contiguous
// slice of the workspace samples, and the grid meets twice:
//   (1) after the forward pair pass, to agree on sum/max of q = q_base + q_iter (renormalize
//       needs both before the importance ratio exists);
//   (2) after the gradient / KL pass, where the last CTA to arrive reduces the per-CTA
//       partials in a fixed order and runs the adjoint sweep.
// With several ranks (samples sharded over GPUs) the same two meeting points carry the
// cross-GPU exchange: the leader CTA stores its rank's totals into every peer's mailbox over
// NVLink (plain st.global on peer-mapped pointers, release flag) and spins on its own mailbox,
// so the collective is a few hundred bytes of P2P stores inside the kernel - no NCCL launch.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstring>

#include "klerg_common.cuh"
#include "klerg_dyn.cuh"
#include "klerg_pair.cuh"

namespace klerg {

// ---- mailbox layout (symmetric across ranks; see klerg_mailbox_bytes) ----------------------
constexpr int MB_MAXW = 8;                       // ranks
// Every value travels as two 8-byte words {32 payload bits, 32-bit tag} (the "LL" scheme of NCCL): a
// naturally aligned 8-byte store is single-copy atomic, so data and flag arrive together and neither
// a system fence nor a separate flag round trip is needed.  tag = mailbox epoch + 1.
constexpr int MB_A_VALS = 2 * FUSED_MAXG;                              // exchange A: {sum, max} per candidate
constexpr int MB_B_VALS = KLERG_MAX_H * KLERG_MAX_D + 2 * FUSED_MAXG;  // exchange B: gradient partials + KL terms
constexpr size_t MB_A_BYTES = (size_t)2 * MB_MAXW * MB_A_VALS * 16;
constexpr size_t MB_B_BYTES = (size_t)2 * MB_MAXW * MB_B_VALS * 16;
constexpr size_t MB_EPOCH_OFF = MB_A_BYTES + MB_B_BYTES;  // u64, written by the owner only
constexpr size_t MB_BYTES = MB_EPOCH_OFF + 64;

__device__ __forceinline__ unsigned long long* mb_a(void* base, int par, int r) {
  return (unsigned long long*)base + (size_t)(par * MB_MAXW + r) * MB_A_VALS * 2;
}
__device__ __forceinline__ unsigned long long* mb_b(void* base, int par, int r) {
  return (unsigned long long*)((char*)base + MB_A_BYTES) + (size_t)(par * MB_MAXW + r) * MB_B_VALS * 2;
}
__device__ __forceinline__ unsigned long long* mb_epoch(void* base) {
  return (unsigned long long*)((char*)base + MB_EPOCH_OFF);
}
__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, unsigned tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long t = (unsigned long long)tag << 32;
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot), "l"((bits & 0xffffffffull) | t) : "memory");
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot + 1), "l"((bits >> 32) | t) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const unsigned long long* slot, unsigned tag, double& v) {
  unsigned long long w0, w1;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w0) : "l"(slot) : "memory");
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w1) : "l"(slot + 1) : "memory");
  if ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) return false;
  v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  return true;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Bounded spins: a meeting point that is never reached (a lost peer, a launch that was not
// co-resident) must not hang the GPU.  After SPIN_LIMIT polls the waiter records a sticky fault
// in ctrl[5] and carries on with whatever is there; klerg_fused_fault() reports it.
constexpr long long SPIN_LIMIT = 1ll << 22;
#define KLERG_SPIN_UNTIL(cond, ctrl)                 \
  for (long long spin_ = 0; !(cond); ++spin_) {      \
    if (spin_ > SPIN_LIMIT) {                        \
      (ctrl)[5] = 1u;                                \
      break;                                         \
    }                                                \
  }

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
#else
#define KLERG_STAMP_DECL
#define KLERG_STAMP(i)
#endif

struct Peers {
  int world, rank;
  void* mail[MB_MAXW];  // mail[r] = rank r's mailbox mapped into this process (mail[rank] = local)
};

struct EvalArgs {
  KernelDev k;
  DynDev d;
  BarDev bar;
  AdjParams ap;
  Peers peers;
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
  double* kl_out;        // [2] {sum p(log p - log c), sum c} over all ranks, or NULL
};

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

// ---------------------------------------------------------------------------
// Meeting point 1: every CTA has written part_tot[blk][g][2]; on return world_tot[g][2]
// (all CTAs, all ranks) is readable by every thread.  The last CTA to arrive is the leader.
// ---------------------------------------------------------------------------
__device__ void meet_totals(const EvalArgs& a, int G, unsigned epoch, unsigned mepoch, int* sh_flag,
                            double* sh_world /* [2G] */) {
  unsigned* ctrl = ws_fused_ctrl(a.ws);
  const unsigned nblk = gridDim.x;
  const double* part = ws_fused_tot(a.ws);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (a.peers.world <= 1) {
    // single GPU: count arrivals, then every CTA combines the per-CTA partials itself (one L2 round trip)
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&ctrl[0], 1u);
      KLERG_SPIN_UNTIL(ld_acquire_u32(&ctrl[0]) >= nblk, ctrl)
    }
    __syncthreads();
    if (warp == 0) {
      for (int g = 0; g < G; ++g) {
        double s = 0.0, m = -INFINITY;
        for (unsigned b = lane; b < nblk; b += 32) {
          s += __ldcg(&part[((size_t)b * FUSED_MAXG + g) * 2 + 0]);
          m = fmax(m, __ldcg(&part[((size_t)b * FUSED_MAXG + g) * 2 + 1]));
        }
        s = warp_reduce(RED_SUM, s);
        m = warp_reduce(RED_MAX, m);
        if (lane == 0) {
          sh_world[2 * g] = s;
          sh_world[2 * g + 1] = m;
        }
      }
    }
    __syncthreads();
    return;
  }
  // several ranks: the last CTA to arrive combines, exchanges with the peers over NVLink and publishes
  const unsigned go_val = 2u * epoch + 1u;
  double* world = ws_fused_world(a.ws);
  const int par = mepoch & 1;
  const unsigned tag = ((mepoch + 1u) << 6) | 63u;
  const int nq = 2 * G;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(&ctrl[0], 1u);
    *sh_flag = (t == nblk - 1);
  }
  __syncthreads();
  if (*sh_flag) {
    __threadfence();
    if (warp == 0) {
      for (int g = 0; g < G; ++g) {
        double s = 0.0, m = -INFINITY;
        for (unsigned b = lane; b < nblk; b += 32) {
          s += __ldcg(&part[((size_t)b * FUSED_MAXG + g) * 2 + 0]);
          m = fmax(m, __ldcg(&part[((size_t)b * FUSED_MAXG + g) * 2 + 1]));
        }
        s = warp_reduce(RED_SUM, s);
        m = warp_reduce(RED_MAX, m);
        // all-gather over NVLink: lane r stores this rank's pair into rank r's mailbox (data + tag per word)
        if (lane < a.peers.world) {
          unsigned long long* slot = mb_a(a.peers.mail[lane], par, a.peers.rank) + 4 * g;
          ll_store(slot, s, tag);
          ll_store(slot + 2, m, tag);
        }
      }
      if (lane < nq) {
        double v = (lane & 1) ? -INFINITY : 0.0;
        for (int r = 0; r < a.peers.world; ++r) {
          const unsigned long long* slot = mb_a(a.peers.mail[a.peers.rank], par, r) + 2 * lane;
          double x = 0.0;
          KLERG_SPIN_UNTIL(ll_try_load(slot, tag, x), ctrl)
          v = (lane & 1) ? fmax(v, x) : v + x;
        }
        world[lane] = v;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release_u32(&ctrl[2], go_val);
    }
  } else if (threadIdx.x == 0) {
    KLERG_SPIN_UNTIL(ld_acquire_u32(&ctrl[2]) == go_val, ctrl)
  }
  __syncthreads();
  if ((int)threadIdx.x < nq) sh_world[threadIdx.x] = __ldcg(&world[threadIdx.x]);
  __syncthreads();
}

// Meeting point 2: returns true (all threads) in the last CTA to arrive.
__device__ bool meet_last(const EvalArgs& a, int* sh_flag, int slot = 1, unsigned round = 0) {
  unsigned* ctrl = ws_fused_ctrl(a.ws);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(&ctrl[slot], 1u);
    *sh_flag = (t == (round + 1u) * gridDim.x - 1u);
  }
  __syncthreads();
  const bool last = *sh_flag != 0;
  if (last) __threadfence();
  return last;
}

// Plain grid barrier on ctrl[slot] (the counter is reset by the CTA that finishes the launch).
__device__ void meet_all(const EvalArgs& a, int slot, unsigned round = 0) {
  unsigned* ctrl = ws_fused_ctrl(a.ws);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&ctrl[slot], 1u);
    KLERG_SPIN_UNTIL(ld_acquire_u32(&ctrl[slot]) >= (round + 1u) * gridDim.x, ctrl)
    __threadfence();
  }
  __syncthreads();
}

// Cross-rank all-gather of `n` doubles held in shared memory (sh_vals) by the last CTA:
// on return sh_vals[i] = sum over ranks (rank order) of the ranks' sh_vals[i].
__device__ void exchange_sum(const EvalArgs& a, unsigned mepoch, double* sh_vals, int n, unsigned round = 0) {
  if (a.peers.world <= 1) return;
  // one use per target: slots alternate with (epoch + target), the tag names both (a rank can run at most one
  // exchange ahead of a peer, because finishing an exchange needs every peer's contribution to it)
  const int par = (mepoch + round) & 1;
  const unsigned tag = ((mepoch + 1u) << 6) | (round & 31u);
  unsigned* ctrl = ws_fused_ctrl(a.ws);
  __syncthreads();
  for (int e = threadIdx.x; e < n * a.peers.world; e += blockDim.x) {
    const int r = e / n, i = e - r * n;
    ll_store(mb_b(a.peers.mail[r], par, a.peers.rank) + 2 * i, sh_vals[i], tag);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double v = 0.0;
    for (int r = 0; r < a.peers.world; ++r) {
      double x = 0.0;
      KLERG_SPIN_UNTIL(ll_try_load(mb_b(a.peers.mail[a.peers.rank], par, r) + 2 * i, tag, x), ctrl)
      v += x;
    }
    sh_vals[i] = v;
  }
  __syncthreads();
}

// contiguous sample slice of this CTA: [lo, hi) with lo % 8 == 0, hi <= ld
__device__ __forceinline__ void cta_slice(int64_t N, int64_t ld, int64_t& lo, int64_t& hi) {
  int64_t per = (N + gridDim.x - 1) / gridDim.x;
  per = (per + 7) & ~(int64_t)7;
  lo = (int64_t)blockIdx.x * per;
  hi = lo + per;
  if (hi > ld) hi = ld;
  if (lo > ld) lo = ld;
}

// Forward pair pass of one trajectory (T duplicated rows at sh_x2) over this CTA's slice:
// v[i] = q_base[i] + inv_nu * sum_t psi; returns the slice's {sum, max} of v over i < N.
template <int D, int P>
__device__ __forceinline__ void forward_slice_p(const EvalArgs& a, const u64* sh_x2, int T, float* v_out, int64_t lo,
                                                int64_t hi, double& tsum, double& tmax) {
  constexpr int SPT = 2 * P;
  for (int64_t i0 = lo + (int64_t)threadIdx.x * SPT; i0 < hi; i0 += (int64_t)blockDim.x * SPT) {
    u64 s2[D][P], acc[P];
    float emin[SPT];
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
    float qb[SPT];
#pragma unroll
    for (int q = 0; q < SPT; ++q) qb[q] = (a.q_base && i0 + q < a.N) ? a.q_base[i0 + q] : 0.f;
#pragma unroll
    for (int q = 0; q < P; ++q) acc[q] = pack2(0.f, 0.f);
    pair_forward<D, P, 0>(sh_x2, T, s2, acc, emin);
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
  }
}

// Samples per thread-iteration: 2 (one packed pair) or 4.  A slice of n samples costs ceil(n / (threads * 2P)) * P
// pair-iterations per thread; for slices of a few samples per thread the quantisation decides (e.g. 6757 samples on
// 512 threads: 4 iterations of two pairs = 8, or 7 iterations of one pair = 7).
__device__ __forceinline__ bool narrow_pairs(int64_t lo, int64_t hi) {
  const int64_t n = hi - lo, bd = blockDim.x;
  const int64_t it1 = (n + bd * 2 - 1) / (bd * 2), it2 = (n + bd * 4 - 1) / (bd * 4);
  return it1 < 2 * it2 || it1 <= 1;
}

template <int D>
__device__ __forceinline__ void forward_slice(const EvalArgs& a, const u64* sh_x2, int T, float* v_out, int64_t lo,
                                              int64_t hi, double& tsum, double& tmax) {
  if (narrow_pairs(lo, hi))
    forward_slice_p<D, 1>(a, sh_x2, T, v_out, lo, hi, tsum, tmax);
  else
    forward_slice_p<D, 2>(a, sh_x2, T, v_out, lo, hi, tsum, tmax);
}

// Forward pair pass of G candidate trajectories over this CTA's slice: samples are loaded once per thread and swept
// against every candidate (the per-candidate totals live in a small indexed array: two local-memory accesses per
// H pairs); v[g][i] to HBM, per-CTA {sum, max} partials per candidate.
template <int D, int P>
__device__ __forceinline__ void forward_candidates(const EvalArgs& a, const u64* s_x2, int G, int H, int64_t lo, int64_t hi,
                                                   double* s_red) {
  constexpr int SPT = 2 * P, DP = Row2<D>::DP;
  const int tid = threadIdx.x;
  double tsum[FUSED_MAXG], tmax[FUSED_MAXG];
  for (int g = 0; g < FUSED_MAXG; ++g) {
    tsum[g] = 0.0;
    tmax[g] = -INFINITY;
  }
  for (int64_t i0 = lo + (int64_t)tid * SPT; i0 < hi; i0 += (int64_t)blockDim.x * SPT) {
    u64 s2[D][P];
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
    float qb[SPT];
#pragma unroll
    for (int q = 0; q < SPT; ++q) qb[q] = (a.q_base && i0 + q < a.N) ? a.q_base[i0 + q] : 0.f;
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      u64 acc[P];
      float emin[SPT], o[SPT];
#pragma unroll
      for (int q = 0; q < P; ++q) acc[q] = pack2(0.f, 0.f);
      pair_forward<D, P, 0>(s_x2 + (size_t)g * H * DP, H, s2, acc, emin);
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
      float* vp = a.v + (size_t)g * a.ld + i0;
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
        double* part = ws_fused_tot(a.ws) + ((size_t)blockIdx.x * FUSED_MAXG + g) * 2;
        part[0] = ra[g];
        part[1] = rb[g];
      }
  }
}

// ---------------------------------------------------------------------------
// shared-memory carve-up
// ---------------------------------------------------------------------------
// Row stride (floats) of the staged sample tiles: a compile-time constant so that the D+2 row addresses of a
// tile are immediates off one base register (no address chain, fewer live registers in the pair loop).
constexpr int TS_ROW = 2048;

struct SmemPlan {
  size_t u, traj, dbarr, P, x2, xs, tile, part, red, misc, total;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

template <int D>
__host__ __device__ inline SmemPlan plan_grad(int H, int S, int A, bool roll, int ts, int nwarps, int WT) {
  SmemPlan p;
  size_t o = 0;
  p.u = o;     o = align16(o + sizeof(float) * H * A);
  p.traj = o;  o = align16(o + sizeof(float) * (H + 1) * S);
  p.dbarr = o; o = align16(o + sizeof(float) * H * S);
  p.P = o;     o = align16(o + (roll ? sizeof(float) * H * A * A : 0));
  p.x2 = o;    o = align16(o + sizeof(u64) * H * Row2<D>::DP);
  p.xs = o;    o = align16(o + sizeof(float) * H * D);
  (void)ts;
  size_t tile = sizeof(float) * (size_t)2 * (D + 2) * TS_ROW;  // two cp.async buffers of rows s_0..s_{D-1}, v->w, p
  const size_t adj = sizeof(double) * ((size_t)H * D + 2) + sizeof(float) * ((size_t)H * S + adjoint_scratch_floats(H, A));
  const size_t rot = roll ? sizeof(float) * rollout_rot_floats(1, H) : 0;
  if (tile < adj) tile = adj;  // the adjoint phase and the ROLL rollout reuse the tile area
  if (tile < rot) tile = rot;
  p.tile = o;  o = align16(o + tile);
  p.part = o;  o = align16(o + sizeof(float) * (size_t)nwarps * WT * D);
  p.red = o;   o = align16(o + sizeof(double) * 32 * 4);
  p.misc = o;  o = align16(o + 16 + sizeof(float) * (KLERG_MAX_S + 9) + 16 + 32);
  p.total = o;
  return p;
}

// ---------------------------------------------------------------------------
// gradient eval
// ---------------------------------------------------------------------------
// MIXED: one chunk per warp, the last a.nwide warps own WT+1 states and the others WT, so that H states tile
// any warp count exactly (no idle state slots) and the warp count can be a multiple of the 4 SM sub-partitions.
template <int D, int WT, int MAXT, bool MIXED>
__global__ void __launch_bounds__(MAXT) eval_grad_kernel(const EvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int WTA = MIXED ? WT + 1 : WT;  // accumulator rows per warp
  const int H = a.H, S = a.d.S, A = a.d.A;
  const bool roll = a.d.kind == KLERG_DYN_ROLL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const SmemPlan sp = plan_grad<D>(H, S, A, roll, a.ts, nwarps, WTA);
  float* s_u = (float*)(smem + sp.u);
  float* s_traj = (float*)(smem + sp.traj);
  float* s_dbarr = (float*)(smem + sp.dbarr);
  float* s_P = roll ? (float*)(smem + sp.P) : nullptr;
  u64* s_x2 = (u64*)(smem + sp.x2);
  float* s_xs = (float*)(smem + sp.xs);
  float* s_tile = (float*)(smem + sp.tile);
  float* s_part = (float*)(smem + sp.part);
  double* s_red = (double*)(smem + sp.red);
  int* s_flag = (int*)(smem + sp.misc);
  unsigned* s_epoch = (unsigned*)(smem + sp.misc) + 1;
  float* s_bsum = (float*)(smem + sp.misc) + 2;
  constexpr int DP = Row2<D>::DP;

  KLERG_STAMP_DECL;
  KLERG_STAMP(0);
  // ---- phase 0: rollout (every CTA) ---------------------------------------------------------------
  float* s_x0 = (float*)(smem + sp.misc) + 4;  // [S] (+ [9] R0)
  unsigned long long* s_bar = (unsigned long long*)(smem + ((sp.misc + 16 + sizeof(float) * (KLERG_MAX_S + 9) + 16 + 7) & ~(size_t)7));  // [2]
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid; e < H * A; e += blockDim.x) s_u[e] = a.u[e];
  if (tid < S) s_x0[tid] = a.x0[tid];
  if (a.R0 && tid >= 32 && tid < 41) s_x0[KLERG_MAX_S + tid - 32] = a.R0[tid - 32];
  if (tid == 0) {
    s_epoch[0] = ws_fused_ctrl(a.ws)[3];
    s_epoch[2] = a.peers.world > 1 ? (unsigned)*mb_epoch(a.peers.mail[a.peers.rank]) : 0u;
  }
  __syncthreads();
  KLERG_STAMP(8);
  rollout_block(a.d, a.bar, s_x0, a.R0 ? s_x0 + KLERG_MAX_S : nullptr, s_u, 1, H, s_traj, s_dbarr, s_P, s_tile, (float*)s_red, s_bsum, nullptr);
  KLERG_STAMP(9);
  const unsigned epoch = s_epoch[0], mepoch = s_epoch[2];
  for (int e = tid; e < H * DP; e += blockDim.x) {
    const int t = e / DP, d = e - t * DP;
    float v = 0.f;
    if (d < D) {
      v = s_traj[t * S + a.k.explr[d]] * a.k.a[d];  // pre-step states (Robot.forward, klerg.py:419-431)
      s_xs[t * D + d] = v;
    }
    s_x2[e] = pack2(v, v);
  }
  if (blockIdx.x == 0 && a.traj)
    for (int e = tid; e < (H + 1) * S; e += blockDim.x) a.traj[e] = s_traj[e];
  __syncthreads();

  KLERG_STAMP(1);
  // ---- phase 1: forward pair pass, slice totals -------------------------------------------------
  int64_t lo, hi;
  cta_slice(a.N, a.ld, lo, hi);
  {
    double tsum = 0.0, tmax = -INFINITY;
    forward_slice<D>(a, s_x2, H, a.v, lo, hi, tsum, tmax);
    const int kinds[2] = {RED_SUM, RED_MAX};
    double vals[2] = {tsum, tmax};
    block_reduce<2>(kinds, vals, s_red);
    if (tid == 0) {
      double* part = ws_fused_tot(a.ws) + (size_t)blockIdx.x * FUSED_MAXG * 2;
      part[0] = vals[0];
      part[1] = vals[1];
    }
  }
  KLERG_STAMP(2);
  // The first sample tile of the gradient pass (samples, this CTA's own v, p) does not depend on the grid-wide
  // totals: its TMA copies are issued now and land while the CTAs meet.
  fence_proxy_async();  // v was written with ordinary stores and is read back by TMA
  __syncthreads();
  if (tid == 0 && hi > lo) {
    const unsigned bytes = 4u * (unsigned)min((int64_t)a.ts, hi - lo);
    mbar_expect_tx(&s_bar[0], (D + 2) * bytes);
#pragma unroll
    for (int d = 0; d < D; ++d) tma_bulk_g2s(s_tile + (size_t)d * TS_ROW, a.packed + (int64_t)d * a.ld + lo, bytes, &s_bar[0]);
    tma_bulk_g2s(s_tile + (size_t)D * TS_ROW, a.v + lo, bytes, &s_bar[0]);
    tma_bulk_g2s(s_tile + (size_t)(D + 1) * TS_ROW, a.p + lo, bytes, &s_bar[0]);
  }
  double* s_world = s_red + 32 * 2;  // [2]
  meet_totals(a, 1, epoch, mepoch, s_flag, s_world);
  KLERG_STAMP(3);
  const double vsum = s_world[0];
  const double vmax = s_world[1];
  if (blockIdx.x == 0 && tid == 0 && a.totals) {
    a.totals[0] = vsum;
    a.totals[1] = vmax;
  }
  const float vsum_f = (float)vsum;  // the reference divides by the fp32 sum
  const float maxc_f = (float)fmax(vmax / vsum, (double)a.floor);

  // ---- phase 2: importance ratio + gradient pair pass ----------------------------------------------
  const int ts = a.ts;
  // loop-invariant launch parameters of the pair loop, pinned through shared memory (see pin_params)
  if (tid == 0) s_flag[4 + KLERG_MAX_S + 9] = a.nsub * 64;
  __syncthreads();
  const int pb_step = ((volatile int*)s_flag)[4 + KLERG_MAX_S + 9];
  const int HD = H * D;
  const int nblk = gridDim.x;
  const int gstride = (nblk + 31) & ~31;  // partial layout [e][gstride]: the final reduce reads rows coalesced
  const bool want_kl = a.kl_out != nullptr || a.cost != nullptr;
  int tile_seq = 0;  // tiles streamed so far in this launch (same in every thread)
  for (int kt = 0; kt < a.K; ++kt) {  // belief targets: the forward pass above is shared, p_k differs
  const float* p_k = a.p + (int64_t)kt * a.p_stride;
  double kl_a = 0.0, kl_c = 0.0;
  for (int r = 0; r < a.rounds; ++r) {
    int cw = warp % a.nchr, sub = warp / a.nchr;
    int t0 = (r * a.nchr + cw) * WT, my_wt = WT;
    bool active = sub < a.nsub && t0 < H;
    if (MIXED) {
      const int narrow = nwarps - a.nwide;
      const bool wide = warp >= narrow;
      t0 = warp * WT + (wide ? warp - narrow : 0);
      my_wt = WT + (wide ? 1 : 0);
      sub = 0;
      active = true;
    }
    u64 xs2[WTA][D], acc[WTA][D];
#pragma unroll
    for (int k = 0; k < WTA; ++k)
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float x = (active && k < my_wt && t0 + k < H) ? s_xs[(t0 + k) * D + d] : 0.f;
        xs2[k][d] = pack2(x, x);
        acc[k][d] = pack2(0.f, 0.f);
      }
    // Sample tiles stream global -> shared with cp.async, one tile ahead of the pair math:
    // rows s_0..s_{D-1} (scaled samples), q_base + q_iter (turned into the importance ratio in place), p.
    const int nt = (int)((hi - lo + ts - 1) / ts);
    // One thread issues the D+2 row copies of a tile as TMA bulk copies (contiguous runs, no descriptors) that
    // complete on the tile buffer's mbarrier; everybody else keeps computing.  Tile g of this launch uses buffer
    // g & 1 and the (g >> 1)-th phase of its barrier.
    auto issue_tile = [&](int k) {
      if (tid == 0) {
        const int gidx = tile_seq + k;
        float* buf = s_tile + (size_t)(gidx & 1) * (D + 2) * TS_ROW;
        const int64_t base = lo + (int64_t)k * ts;
        const unsigned bytes = 4u * (unsigned)min((int64_t)ts, hi - base);  // multiple of 16
        fence_proxy_async();  // the buffer was last written with ordinary stores (importance ratio, padding)
        mbar_expect_tx(&s_bar[gidx & 1], (D + 2) * bytes);
#pragma unroll
        for (int d = 0; d < D; ++d)
          tma_bulk_g2s(buf + (size_t)d * TS_ROW, a.packed + (int64_t)d * a.ld + base, bytes, &s_bar[gidx & 1]);
        tma_bulk_g2s(buf + (size_t)D * TS_ROW, a.v + base, bytes, &s_bar[gidx & 1]);
        tma_bulk_g2s(buf + (size_t)(D + 1) * TS_ROW, p_k + base, bytes, &s_bar[gidx & 1]);
      }
    };
    if (nt > 0 && tile_seq > 0) issue_tile(0);  // the very first tile of the launch was issued before the meeting point
    for (int k = 0; k < nt; ++k) {
      const int gidx = tile_seq + k;
      float* buf = s_tile + (size_t)(gidx & 1) * (D + 2) * TS_ROW;
      const int64_t base = lo + (int64_t)k * ts;
      const int cnt = (int)min((int64_t)ts, hi - base);
      const int cnt64 = (cnt + 63) & ~63;
      float* wrow = buf + (size_t)D * TS_ROW;
      const float* prow = buf + (size_t)(D + 1) * TS_ROW;
      {
        unsigned* ctrl = ws_fused_ctrl(a.ws);
        KLERG_SPIN_UNTIL(mbar_try_wait(&s_bar[gidx & 1], (unsigned)(gidx >> 1) & 1u), ctrl)
      }
      for (int c = tid; c < (cnt64 >> 2); c += blockDim.x) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = (c << 2) + q;
          const int64_t i = base + e;
          float w = 0.f;
          if (e < cnt && i < a.N) {
            const float cc = fmaxf(__fdividef(wrow[e], vsum_f), a.floor);
            const float pi = prow[e];
            w = __fdividef(pi * maxc_f, cc);  // p/q with q = c / max c  (klerg.py:436)
            if (want_kl && r == 0) {
              kl_a += (double)(pi * (logf(pi) - logf(cc)));
              kl_c += (double)cc;
            }
          } else if (e >= cnt) {
#pragma unroll
            for (int d = 0; d < D; ++d) buf[(size_t)d * TS_ROW + e] = 0.f;
          }
          wrow[e] = w;
        }
      }
      __syncthreads();  // tile k is ready for everyone; everyone is done with tile k-1
      if (k + 1 < nt) issue_tile(k + 1);
      if (active) {
        for (int pb = sub * 64; pb < cnt64; pb += pb_step) {
          const int i = pb + 2 * lane;
          u64 s2[D];
#pragma unroll
          for (int d = 0; d < D; ++d) s2[d] = *reinterpret_cast<const u64*>(&buf[(size_t)d * TS_ROW + i]);
          const u64 w2 = *reinterpret_cast<const u64*>(&wrow[i]);
          pair_gradient<D, WTA>(xs2, s2, w2, acc, my_wt == WTA);
        }
      }
    }
    tile_seq += nt;
    __syncthreads();
    // lanes -> warp sums -> CTA partial for this round's states
#pragma unroll
    for (int k = 0; k < WTA; ++k)
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float x, y;
        unpack2(acc[k][d], x, y);
        const float v = warp_sum_f(x + y);
        if (lane == 0) s_part[(warp * WTA + k) * D + d] = v;
      }
    __syncthreads();
    float* gpart = (float*)ws_fused_grad(a.ws);  // [H*D][gstride] fp32 (the CTA sums are fp32 values)
    if (MIXED) {
      const int narrow = nwarps - a.nwide, narrow_states = narrow * WT;
      for (int e = tid; e < HD; e += blockDim.x) {
        const int t = e / D, d = e - t * D;
        int w, k;
        if (t < narrow_states) {
          w = t / WT;
          k = t - w * WT;
        } else {
          const int tt = t - narrow_states;
          w = narrow + tt / (WT + 1);
          k = tt - (w - narrow) * (WT + 1);
        }
        gpart[(size_t)e * gstride + blockIdx.x] = s_part[(w * WTA + k) * D + d];
      }
    } else
    for (int e = tid; e < a.nchr * WT * D; e += blockDim.x) {
      const int c = e / (WT * D), kd = e - c * (WT * D);
      const int t = (r * a.nchr + c) * WT + kd / D;
      if (t < H) {
        float v = 0.f;
        for (int sb = 0; sb < a.nsub; ++sb) v += s_part[((sb * a.nchr + c) * WT) * D + kd];
        gpart[(size_t)(t * D + kd % D) * gstride + blockIdx.x] = v;
      }
    }
  }
  if (want_kl) {
    const int kinds[2] = {RED_SUM, RED_SUM};
    double vals[2] = {kl_a, kl_c};
    block_reduce<2>(kinds, vals, s_red);
    if (tid == 0) {
      double* part = ws_fused_kl(a.ws) + (size_t)blockIdx.x * FUSED_MAXG * 2 + (kt & 1) * 2;
      part[0] = vals[0];
      part[1] = vals[1];
    }
  }

  // ---- phase 3: every CTA reduces a few gradient entries over all CTA partials (fixed order), then the last
  //      CTA to finish collects the H*D sums, exchanges them with the peers and runs the adjoint ----------------
  KLERG_STAMP(4);
  meet_all(a, 1, kt);
  double* gfin = (double*)((char*)ws_fused_grad(a.ws) + FUSED_GRAD / 2);  // [H*D]
  {
    const float* gpart = (const float*)ws_fused_grad(a.ws);
    for (int e = blockIdx.x * nwarps + warp; e < HD; e += nblk * nwarps) {
      const float* row = gpart + (size_t)e * gstride;
      double v = 0.0;
      for (int b = lane; b < nblk; b += 32) v += (double)__ldcg(row + b);
      v = warp_reduce(RED_SUM, v);
      if (lane == 0) gfin[e] = v;
    }
  }
  if (!meet_last(a, s_flag, 4, kt)) continue;
  KLERG_STAMP(5);
  double* s_val = (double*)s_tile;               // [HD + 2]
  float* s_g = (float*)(s_val + HD + 2);         // [H][S]
  float* s_scr = s_g + H * S;                    // adjoint scratch
  for (int e = tid; e < HD; e += blockDim.x) s_val[e] = __ldcg(&gfin[e]);
  if (want_kl) {
    const double* klp = ws_fused_kl(a.ws);
    if (warp == 0) {
      double s0 = 0.0, s1 = 0.0;
      for (int b = lane; b < nblk; b += 32) {
        s0 += __ldcg(&klp[(size_t)b * FUSED_MAXG * 2 + (kt & 1) * 2 + 0]);
        s1 += __ldcg(&klp[(size_t)b * FUSED_MAXG * 2 + (kt & 1) * 2 + 1]);
      }
      s0 = warp_reduce(RED_SUM, s0);
      s1 = warp_reduce(RED_SUM, s1);
      if (lane == 0) {
        s_val[HD] = s0;
        s_val[HD + 1] = s1;
      }
    }
  } else if (tid == 0) {
    s_val[HD] = 0.0;
    s_val[HD + 1] = 1.0;
  }
  __syncthreads();
  KLERG_STAMP(6);
  exchange_sum(a, mepoch, s_val, HD + 2, kt);
  for (int e = tid; e < H * S; e += blockDim.x) s_g[e] = 0.f;
  __syncthreads();
  for (int e = tid; e < HD; e += blockDim.x) {
    const int d = e % D;
    s_g[(e / D) * S + a.k.explr[d]] = (float)(s_val[e] * (double)a.k.gfac[d]);
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
    const double sa = s_val[HD], sc = s_val[HD + 1];
    if (a.kl_out) {
      a.kl_out[2 * kt] = sa;
      a.kl_out[2 * kt + 1] = sc;
    }
    if (a.cost) {
      const double spv = a.p_stats[kt];
      // KL of the PRE-step footprint (what backward() differentiates) + barrier of the post-step states
      a.cost[kt] = (float)(sa / spv - log(spv) + log(sc)) + *s_bsum;
    }
    unsigned* ctrl = ws_fused_ctrl(a.ws);
    if (kt == a.K - 1) {  // the CTA that finishes the last target closes the launch
      ctrl[0] = 0;
      ctrl[1] = 0;
      ctrl[4] = 0;
      ctrl[3] = epoch + 1;
      if (a.peers.world > 1) *mb_epoch(a.peers.mail[a.peers.rank]) = (unsigned long long)mepoch + 1ull;
    }
#ifdef KLERG_STAMPS
    // phase stamps of the CTA that finished last (SM cycles since its start): profiling aid
    KLERG_STAMP(7);
    long long* dbg = (long long*)(ctrl + 16);
    for (int i = 0; i < 10; ++i) dbg[i] = stamp[i] - stamp[0];
    for (int i = 0; i < 8; ++i) dbg[10 + i] = g_ro_stamp[i] - g_ro_stamp[0];
#endif
  }
  __syncthreads();  // the CTA that ran the adjoint reuses its tile area for the next target
  }  // targets
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
  p.tile = o; o = align16(o + (roll ? sizeof(float) * rollout_rot_floats(G, H) : 0));
  p.red = o;  o = align16(o + sizeof(double) * 32 * 2 * FUSED_MAXG);
  p.misc = o; o = align16(o + 64 + sizeof(float) * FUSED_MAXG);
  p.total = o;
  return p;
}

template <int D>
__global__ void __launch_bounds__(512) eval_cost_kernel(const EvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int H = a.H, S = a.d.S, A = a.d.A, G = a.G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const SmemPlan sp = plan_cost<D>(G, H, S, A, a.d.kind == KLERG_DYN_ROLL);
  float* s_u = (float*)(smem + sp.u);
  float* s_traj = (float*)(smem + sp.traj);
  u64* s_x2 = (u64*)(smem + sp.x2);
  double* s_red = (double*)(smem + sp.red);
  int* s_flag = (int*)(smem + sp.misc);
  unsigned* s_epoch = (unsigned*)(smem + sp.misc) + 1;
  float* s_bsum = (float*)(smem + sp.misc + 64);
  constexpr int DP = Row2<D>::DP;

  for (int e = tid; e < G * H * A; e += blockDim.x) s_u[e] = a.u[e];
  if (tid == 0) {
    s_epoch[0] = ws_fused_ctrl(a.ws)[3];
    s_epoch[2] = a.peers.world > 1 ? (unsigned)*mb_epoch(a.peers.mail[a.peers.rank]) : 0u;
  }
  __syncthreads();
  rollout_block(a.d, a.bar, a.x0, a.R0, s_u, G, H, s_traj, nullptr, nullptr, (float*)(smem + sp.tile), (float*)s_red,
                s_bsum, nullptr);
  const unsigned epoch = s_epoch[0], mepoch = s_epoch[2];
  for (int e = tid; e < G * H * DP; e += blockDim.x) {
    const int g = e / (H * DP), r = e - g * (H * DP);
    const int t = r / DP, d = r - t * DP;
    float v = 0.f;
    if (d < D) v = s_traj[((size_t)g * (H + 1) + t + 1) * S + a.k.explr[d]] * a.k.a[d];  // post-step states (klerg.py:688-691)
    s_x2[e] = pack2(v, v);
  }
  if (blockIdx.x == 0 && a.traj)
    for (int e = tid; e < G * (H + 1) * S; e += blockDim.x) a.traj[e] = s_traj[e];
  __syncthreads();

  int64_t lo, hi;
  cta_slice(a.N, a.ld, lo, hi);
  if (narrow_pairs(lo, hi))
    forward_candidates<D, 1>(a, s_x2, G, H, lo, hi, s_red);
  else
    forward_candidates<D, 2>(a, s_x2, G, H, lo, hi, s_red);
  double* s_world = s_red + 32 * 2;  // [2G]
  meet_totals(a, G, epoch, mepoch, s_flag, s_world);
  if (blockIdx.x == 0 && tid < 2 * G && a.totals) a.totals[tid] = s_world[tid];

  // KL partials: sum_i p_i (log p_i - log c_i), sum_i c_i   (klerg.py:694-699 in closed form)
  // Four samples per thread-iteration (128-bit loads of v and p), reciprocal of the normaliser and lg2-based
  // logarithms: the pass is instruction-bound (IEEE division + logf cost ~60 instructions per sample and candidate,
  // this form ~12); the cost changes by < 1e-6 relative, far inside the 1e-4 parity tolerance.
  float rvs[FUSED_MAXG], maxc[FUSED_MAXG];
  double sa[FUSED_MAXG], sc[FUSED_MAXG];
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
        const float4 t = __ldcg(reinterpret_cast<const float4*>(a.v + (size_t)g * a.ld + i0));
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
  block_reduce_pairs(G, RED_SUM, sa, sc, s_red);
  if (tid == 0) {
#pragma unroll
    for (int g = 0; g < FUSED_MAXG; ++g)
      if (g < G) {
        double* part = ws_fused_kl(a.ws) + ((size_t)blockIdx.x * FUSED_MAXG + g) * 2;
        part[0] = sa[g];
        part[1] = sc[g];
      }
  }

  if (!meet_last(a, s_flag)) return;
  double* s_val = s_red;  // [2G]
  {
    const double* klp = ws_fused_kl(a.ws);
    const int nblk = gridDim.x;
    const int n4 = 2 * G * 4;
    for (int idx = tid; idx < ((n4 + 31) & ~31); idx += blockDim.x) {  // whole warps take part in the shuffles
      const int e = idx >> 2, part = idx & 3;
      double v = 0.0;
      if (idx < n4) {
#pragma unroll 8
        for (int b = part; b < nblk; b += 4) v += __ldcg(&klp[(size_t)b * FUSED_MAXG * 2 + e]);
      }
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (part == 0 && idx < n4) s_val[e] = v;
    }
  }
  __syncthreads();
  exchange_sum(a, mepoch, s_val, 2 * G);
  if (tid < G) {
    const double spv = a.p_stats[0];
    const double dkl = s_val[2 * tid] / spv - log(spv) + log(s_val[2 * tid + 1]);
    a.cost[tid] = (float)dkl + s_bsum[tid];
  }
  if (tid == 0) {
    unsigned* ctrl = ws_fused_ctrl(a.ws);
    ctrl[0] = 0;
    ctrl[1] = 0;
    ctrl[3] = epoch + 1;
    if (a.peers.world > 1) *mb_epoch(a.peers.mail[a.peers.rank]) = (unsigned long long)mepoch + 1ull;
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct GradSchedule {
  int wt, nwarps, nchr, nsub, rounds;
  double eff;
  bool mixed;
  int nwide;
};

static int grad_max_warps(int D) { return D <= 3 ? 20 : (D == 4 ? 16 : 17); }

// Choose states-per-warp WT and the warp grid so that (states x sample sub-streams) tiles the
// CTA's warps with as few idle slots as possible.
static GradSchedule plan_schedule(int D, int H) {
  if (D >= 4) {
    // mixed schedule: 16 warps (4 per SM sub-partition, 128 registers), H = q*16 + r -> r warps own q+1 states.
    // Needs q >= 2 (fewer states per warp would re-read the staged samples too often for the shared-memory bandwidth).
    const int nw = 16, q = H / nw, r = H % nw, wtmax = D == 4 ? 5 : 4;
    if (q >= 2 && r > 0 && q + 1 <= wtmax) {
      GradSchedule m{};
      m.wt = q; m.nwarps = nw; m.nchr = nw; m.nsub = 1; m.rounds = 1; m.eff = 1.0; m.mixed = true; m.nwide = r;
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

template <typename K>
static int coop_launch(K kernel, int nblk, int nthreads, size_t smem, const EvalArgs& a, cudaStream_t stream,
                       const char* what) {
  EvalArgs args = a;
  void* pargs[1] = {(void*)&args};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)nblk), dim3((unsigned)nthreads), pargs,
                                              smem, stream);
  if (e != cudaSuccess) {
    set_error("%s: cooperative launch failed (grid %d x %d threads, %zu B smem): %s", what, nblk, nthreads, smem,
              cudaGetErrorString(e));
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

static int pick_grid(int64_t N, int per_sm, int min_samples_per_cta) {
  int64_t nblk = (N + min_samples_per_cta - 1) / min_samples_per_cta;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (nblk > cap) nblk = cap;
  if (nblk > FUSED_MAXBLK) nblk = FUSED_MAXBLK;
  if (nblk < 1) nblk = 1;
  return (int)nblk;
}

template <int D, int WT, bool MIXED = false>
static int launch_grad_wt(EvalArgs& a, const GradSchedule& s, cudaStream_t stream) {
  constexpr int MAXT = MIXED ? 512 : (D <= 3 ? 20 : (D == 4 ? 16 : 17)) * 32;
  auto kernel = eval_grad_kernel<D, WT, MAXT, MIXED>;
  const int nthreads = s.nwarps * 32;
  const bool roll = a.d.kind == KLERG_DYN_ROLL;
  // tile: up to 2048 samples, but no more than one CTA's slice at full grid
  int64_t per = (a.N + sm_count() - 1) / sm_count();
  int ts = 2048;
  while (ts > 64 && ts / 2 >= per) ts /= 2;
  a.ts = ts;
  a.nchr = s.nchr; a.nsub = s.nsub; a.rounds = s.rounds; a.nwide = s.nwide;
  const SmemPlan sp = plan_grad<D>(a.H, a.d.S, a.d.A, roll, ts, s.nwarps, MIXED ? WT + 1 : WT);
  if (sp.total > 220 * 1024) { set_error("eval_gradient: horizon too long for shared-memory staging"); return -1; }
  const int per_sm = resident_ctas(kernel, nthreads, sp.total);
  if (per_sm < 1) { set_error("eval_gradient: kernel does not fit on an SM (threads=%d smem=%zu)", nthreads, sp.total); return -4; }
  const int nblk = pick_grid(a.N, 1, 128);
  return coop_launch(kernel, nblk, nthreads, sp.total, a, stream, "eval_grad_kernel");
}

template <int D>
static int launch_grad_d(EvalArgs& a, cudaStream_t stream) {
  const GradSchedule s = plan_schedule(D, a.H);
  if constexpr (D >= 4) {
    if (s.mixed) {
      if (s.wt == 2) return launch_grad_wt<D, 2, true>(a, s, stream);
      if (s.wt == 3) return launch_grad_wt<D, 3, true>(a, s, stream);
      if constexpr (D == 4) {
        if (s.wt == 4) return launch_grad_wt<D, 4, true>(a, s, stream);
      }
      set_error("eval_gradient: no mixed schedule for D=%d H=%d", D, a.H);
      return -2;
    }
  }
  switch (s.wt) {
    case 1: return launch_grad_wt<D, 1>(a, s, stream);
    case 2: return launch_grad_wt<D, 2>(a, s, stream);
    case 3: return launch_grad_wt<D, 3>(a, s, stream);
    case 4:
      if constexpr (D <= 4) return launch_grad_wt<D, 4>(a, s, stream);
      break;
    case 5:
      if constexpr (D <= 4) return launch_grad_wt<D, 5>(a, s, stream);
  }
  set_error("eval_gradient: no schedule for D=%d H=%d", D, a.H);
  return -2;
}

template <int D>
static int launch_cost_d(EvalArgs& a, cudaStream_t stream) {
  auto kernel = eval_cost_kernel<D>;
  const SmemPlan sp = plan_cost<D>(a.G, a.H, a.d.S, a.d.A, a.d.kind == KLERG_DYN_ROLL);
  if (sp.total > 200 * 1024) { set_error("eval_costs: G*H too large for shared-memory staging"); return -1; }
  int nthreads = 512;
  const int per_sm = resident_ctas(kernel, nthreads, sp.total);
  if (per_sm < 1) { set_error("eval_costs: kernel does not fit on an SM"); return -4; }
  const int nblk = pick_grid(a.N, 1, 256);
  return coop_launch(kernel, nblk, nthreads, sp.total, a, stream, "eval_cost_kernel");
}

static bool fill_common(EvalArgs& a, const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                        const klerg_peers* peers) {
  if (!make_kernel_dev(k, a.k) || !make_dyn(dyn, a.d) || !make_bar(bar, a.bar)) return false;
  a.peers.world = 1;
  a.peers.rank = 0;
  for (int r = 0; r < MB_MAXW; ++r) a.peers.mail[r] = nullptr;
  if (peers && peers->world > 1) {
    if (peers->world > MB_MAXW || peers->rank < 0 || peers->rank >= peers->world) { set_error("peers: world/rank out of range"); return false; }
    a.peers.world = peers->world;
    a.peers.rank = peers->rank;
    for (int r = 0; r < peers->world; ++r) {
      if (!peers->mailbox[r]) { set_error("peers: mailbox[%d] is null", r); return false; }
      a.peers.mail[r] = peers->mailbox[r];
    }
  }
  return true;
}

}  // namespace klerg

using namespace klerg;

extern "C" size_t klerg_mailbox_bytes(void) { return MB_BYTES; }

// Mailboxes are plain cudaMalloc allocations shared between the one-process-per-GPU ranks with CUDA IPC.
extern "C" int klerg_mailbox_create(void** ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is exchanged as 64 bytes");
  if (!ptr || !handle64) { set_error("mailbox_create: null argument"); return -1; }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, MB_BYTES);
  if (e == cudaSuccess) e = cudaMemset(p, 0, MB_BYTES);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("mailbox_create: %s", cudaGetErrorString(e));
    if (p) cudaFree(p);
    cudaGetLastError();
    return -4;
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

extern "C" int klerg_mailbox_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) { set_error("mailbox_open: null argument"); return -1; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("mailbox_open: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return -4;
  }
  *ptr = p;
  return 0;
}

extern "C" int klerg_mailbox_close(void* ptr, int owner) {
  if (!ptr) return 0;
  cudaError_t e = owner ? cudaFree(ptr) : cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    set_error("mailbox_close: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return -4;
  }
  return 0;
}
extern "C" size_t klerg_fused_fault_offset(void) { return HEAD_COUNTERS + HEAD_MISC + HEAD_GRAD + 5 * sizeof(unsigned); }
extern "C" size_t klerg_debug_stamps_offset(void) { return HEAD_COUNTERS + HEAD_MISC + HEAD_GRAD + 64; }

extern "C" int klerg_eval_gradient_targets(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn,
                                           const klerg_barrier_spec* bar, const klerg_peers* peers, const float* x0,
                                           const float* R0, const float* u, int64_t H, const float* packed, int64_t N,
                                           int64_t ld, const float* q_base, const float* p, int64_t K, int64_t p_stride,
                                           const double* p_stats, float floor, const float* Rinv_diag, float alpha,
                                           const float* ctrl_lo, const float* ctrl_hi, float* v_scratch, float* traj,
                                           double* totals, float* cost, float* dgdx, float* du, float* djdlam,
                                           float* u_star, double* kl_out, void* workspace, void* stream) {
  EvalArgs a{};
  if (K < 1 || K > 32) { set_error("eval_gradient: K must be in 1..32 targets per launch"); return -1; }
  if (K > 1 && (p_stride < N || (p_stride & 3))) { set_error("eval_gradient: p_stride must be >= N and a multiple of 4"); return -1; }
  if (!fill_common(a, k, dyn, bar, peers)) return -1;
  if (H < 1 || H > KLERG_MAX_H) { set_error("eval_gradient: H out of range"); return -1; }
  if (N < 1 || ld < N || (ld & 3)) { set_error("eval_gradient: bad sample sizes"); return -1; }
  if (!workspace || !v_scratch || !dgdx || !du || !djdlam || !u_star) { set_error("eval_gradient: null output/workspace"); return -1; }
  if (((uintptr_t)packed | (uintptr_t)p | (uintptr_t)v_scratch) & 15) { set_error("eval_gradient: packed, p and v_scratch must be 16-byte aligned"); return -1; }
  for (int i = 0; i < a.d.A; ++i) { a.ap.rinv[i] = Rinv_diag[i]; a.ap.clo[i] = ctrl_lo[i]; a.ap.chi[i] = ctrl_hi[i]; }
  a.ap.alpha = alpha;
  a.x0 = x0; a.R0 = R0; a.u = u; a.G = 1; a.H = (int)H; a.packed = packed; a.N = N; a.ld = ld; a.q_base = q_base; a.p = p;
  a.K = (int)K; a.p_stride = p_stride;
  a.p_stats = p_stats; a.floor = floor; a.v = v_scratch; a.ws = workspace; a.traj = traj; a.totals = totals; a.cost = cost;
  a.dgdx = dgdx; a.du = du; a.djdlam = djdlam; a.u_star = u_star; a.kl_out = kl_out;
  switch (a.k.D) {
    case 1: return launch_grad_d<1>(a, (cudaStream_t)stream);
    case 2: return launch_grad_d<2>(a, (cudaStream_t)stream);
    case 3: return launch_grad_d<3>(a, (cudaStream_t)stream);
    case 4: return launch_grad_d<4>(a, (cudaStream_t)stream);
    case 5: return launch_grad_d<5>(a, (cudaStream_t)stream);
    case 6: return launch_grad_d<6>(a, (cudaStream_t)stream);
    default: set_error("eval_gradient: D=%d not instantiated (1..6)", a.k.D); return -2;
  }
}

extern "C" int klerg_eval_gradient(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                                   const klerg_peers* peers, const float* x0, const float* R0, const float* u, int64_t H,
                                   const float* packed, int64_t N, int64_t ld, const float* q_base, const float* p,
                                   const double* p_stats, float floor, const float* Rinv_diag, float alpha,
                                   const float* ctrl_lo, const float* ctrl_hi, float* v_scratch, float* traj,
                                   double* totals, float* cost, float* dgdx, float* du, float* djdlam, float* u_star,
                                   double* kl_out, void* workspace, void* stream) {
  return klerg_eval_gradient_targets(k, dyn, bar, peers, x0, R0, u, H, packed, N, ld, q_base, p, 1, 0, p_stats, floor,
                                     Rinv_diag, alpha, ctrl_lo, ctrl_hi, v_scratch, traj, totals, cost, dgdx, du, djdlam,
                                     u_star, kl_out, workspace, stream);
}

extern "C" int klerg_eval_costs(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                                const klerg_peers* peers, const float* x0, const float* R0, const float* u, int64_t G,
                                int64_t H, const float* packed, int64_t N, int64_t ld, const float* q_base,
                                const float* p, const double* p_stats, float floor, float* v_scratch, float* traj,
                                double* totals, float* cost, void* workspace, void* stream) {
  EvalArgs a{};
  if (!fill_common(a, k, dyn, bar, peers)) return -1;
  if (H < 1 || H > KLERG_MAX_H) { set_error("eval_costs: H out of range"); return -1; }
  if (G < 1 || G > FUSED_MAXG) { set_error("eval_costs: G must be in 1..%d", FUSED_MAXG); return -1; }
  if (N < 1 || ld < N || (ld & 3)) { set_error("eval_costs: bad sample sizes"); return -1; }
  if (!workspace || !v_scratch || !cost) { set_error("eval_costs: null output/workspace"); return -1; }
  a.x0 = x0; a.R0 = R0; a.u = u; a.G = (int)G; a.H = (int)H; a.packed = packed; a.N = N; a.ld = ld; a.q_base = q_base;
  a.p = p; a.p_stats = p_stats; a.floor = floor; a.v = v_scratch; a.ws = workspace; a.traj = traj; a.totals = totals;
  a.cost = cost; a.K = 1; a.p_stride = 0;
  switch (a.k.D) {
    case 1: return launch_cost_d<1>(a, (cudaStream_t)stream);
    case 2: return launch_cost_d<2>(a, (cudaStream_t)stream);
    case 3: return launch_cost_d<3>(a, (cudaStream_t)stream);
    case 4: return launch_cost_d<4>(a, (cudaStream_t)stream);
    case 5: return launch_cost_d<5>(a, (cudaStream_t)stream);
    case 6: return launch_cost_d<6>(a, (cudaStream_t)stream);
    default: set_error("eval_costs: D=%d not instantiated (1..6)", a.k.D); return -2;
  }
}

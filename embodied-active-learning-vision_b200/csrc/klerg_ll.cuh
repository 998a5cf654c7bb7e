// Flag-free exchanges between the CTAs of a fused eval and between the ranks of a sample-sharded
// workspace ("LL" protocol: 8-byte words that carry 32 payload bits and a 32-bit tag).
//
// The fused evals need two kinds of meeting:
//   * all-reduce of a few doubles per CTA ({sum, max} of q per candidate) whose result every CTA of every
//     rank needs before it can go on                              -> ll_allreduce
//   * a many-to-one gather (gradient partials, KL terms) that only the CTA running the adjoint needs
//                                                                 -> slots polled by that CTA (klerg_fused.cu)
// Both are built from the same primitive: the producer stores {payload, tag} words - into its own GPU's
// mailbox and, with plain stores over NVLink, into the peers' mailboxes - and a consumer polls the slot until
// the tag of the current exchange appears.  There are no counters to reset and no leader: a meeting costs one
// store and one (local L2) poll, and every consumer combines the slots in the same fixed order, so all CTAs
// and all ranks hold bit-identical results (the host control flow of the ranks stays in lockstep).
#pragma once
#include "klerg_common.cuh"

namespace klerg {

typedef unsigned long long u64;

struct Peers {
  int world, rank;
  void* mail[MB_MAXW];  // mail[r] = rank r's mailbox mapped into this process (mail[rank] = local)
};

__device__ __forceinline__ unsigned* mb_hdr(void* base) { return (unsigned*)base; }
__device__ __forceinline__ u64* mb_x1(void* base, int par, int r, int b) {
  return (u64*)((char*)base + MB_OFF_X1) + ((size_t)(par * MB_MAXW + r) * LL_MAXBLK + b) * 4;
}
__device__ __forceinline__ u64* mb_l1(void* base, int par, int b) {
  return (u64*)((char*)base + MB_OFF_L1) + ((size_t)par * LL_MAXBLK + b) * LL_MAXNV * 2;
}
__device__ __forceinline__ u64* mb_r1(void* base, int par, int r) {
  return (u64*)((char*)base + MB_OFF_R1) + ((size_t)par * MB_MAXW + r) * LL_MAXNV * 2;
}
__device__ __forceinline__ u64* mb_kl(void* base, int par, int b) {
  return (u64*)((char*)base + MB_OFF_KL) + ((size_t)par * LL_MAXBLK + b) * LL_MAXNV * 2;
}
__device__ __forceinline__ u64* mb_gb(void* base, int par, int r) {
  return (u64*)((char*)base + MB_OFF_GB) + ((size_t)par * MB_MAXW + r) * (LL_MAXHD + 16) * 2;
}
__device__ __forceinline__ u64* mb_gp(void* base, int par, int e) {
  return (u64*)((char*)base + MB_OFF_GP) + ((size_t)par * LL_MAXHD + e) * LL_MAXBLK;
}

__device__ __forceinline__ float* mb_ro(void* base, int par) {
  return (float*)((char*)base + MB_OFF_RO + (size_t)par * (MB_RO / 2));
}

// tag of exchange `id` of the launch / exchange numbered `counter` (never 0: mailboxes start zero-filled)
__device__ __forceinline__ unsigned ll_tag(unsigned counter, unsigned id) {
  return (((counter + 1u) & 0xFFFFFFu) << 8) | (id & 0xFFu);
}

__device__ __forceinline__ void ll_store(u64* slot, double v, unsigned tag) {
  const u64 bits = (u64)__double_as_longlong(v);
  const u64 t = (u64)tag << 32;
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot), "l"((bits & 0xffffffffull) | t) : "memory");
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot + 1), "l"((bits >> 32) | t) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const u64* slot, unsigned tag, double& v) {
  u64 w0, w1;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  if ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) return false;
  v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  return true;
}
// one fp32 value + tag in a single word
__device__ __forceinline__ void ll_store_f32(u64* slot, float v, unsigned tag) {
  const u64 w = ((u64)tag << 32) | (u64)__float_as_uint(v);
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot), "l"(w) : "memory");
}
__device__ __forceinline__ bool ll_try_load_f32(const u64* slot, unsigned tag, float& v) {
  u64 w;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(slot) : "memory");
  if ((unsigned)(w >> 32) != tag) return false;
  v = __uint_as_float((unsigned)(w & 0xffffffffull));
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

// Bounded spins: an exchange that never completes (a lost peer, a launch that was not co-resident) must not
// hang the GPU.  After SPIN_LIMIT polls the waiter records a sticky fault in ctrl[5] and carries on with
// whatever is there; klerg_fused_fault() reports it and the Python planner raises.
constexpr long long SPIN_LIMIT = 1ll << 22;
#define KLERG_SPIN_UNTIL(cond, ctrl)                 \
  for (long long spin_ = 0; !(cond); ++spin_) {      \
    if (spin_ > SPIN_LIMIT) {                        \
      (ctrl)[5] = 1u;                                \
      break;                                         \
    }                                                \
  }

// Programmatic dependent launch (sm_90+): both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Reduce `count` staged records of nv doubles each (s_buf[j * nv + i]) per value i in a fixed order: value i is a
// maximum where bit i of max_mask is set, else a sum.  Result in s_out[i] (all threads may read it after the
// trailing barrier).  s_part: 32 doubles of scratch.  Called by the whole CTA.
__device__ __forceinline__ void ll_reduce_staged(const double* s_buf, int count, int nv, unsigned max_mask, double* s_out,
                                                 double* s_part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (nv <= nwarps) {
    const int wpv = nwarps / nv;  // warps per value
    const int i = warp / wpv, sub = warp - i * wpv;
    if (i < nv) {
      const bool is_max = (max_mask >> i) & 1u;
      double v = is_max ? -INFINITY : 0.0;
      for (int j = sub * 32 + lane; j < count; j += wpv * 32) {
        const double x = s_buf[(size_t)j * nv + i];
        v = is_max ? fmax(v, x) : v + x;
      }
      v = warp_reduce(is_max ? RED_MAX : RED_SUM, v);
      if (lane == 0) s_part[warp] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < nv) {
      const int k = threadIdx.x;
      const bool is_max = (max_mask >> k) & 1u;
      double v = s_part[k * wpv];
      for (int s = 1; s < wpv; ++s) v = is_max ? fmax(v, s_part[k * wpv + s]) : v + s_part[k * wpv + s];
      s_out[k] = v;
    }
  } else {
    for (int i = warp; i < nv; i += nwarps) {
      const bool is_max = (max_mask >> i) & 1u;
      double v = is_max ? -INFINITY : 0.0;
      for (int j = lane; j < count; j += 32) {
        const double x = s_buf[(size_t)j * nv + i];
        v = is_max ? fmax(v, x) : v + x;
      }
      v = warp_reduce(is_max ? RED_MAX : RED_SUM, v);
      if (lane == 0) s_out[i] = v;
    }
  }
  __syncthreads();
}

// All-reduce of nv <= LL_MAXNV doubles per CTA over the vnblk CTAs of each of P.world ranks.
//   s_in [nv]  this CTA's contribution (shared memory, complete before the call)
//   s_out[nv]  result, identical bits in every CTA of every rank
//   s_buf      LL_BUF_VALS + 32 doubles of shared scratch
// One hop when a rank's CTAs can poll every other CTA directly (single GPU, or {sum, max} pairs across ranks);
// otherwise a local stage followed by a rank stage.  `epoch` selects the slot parity and the tags; at most one
// all-reduce per launch may use the same `id`.
__device__ __forceinline__ void ll_allreduce(const Peers& P, int vblk, int vnblk, unsigned epoch, unsigned id, int nv,
                                             unsigned max_mask, const double* s_in, double* s_out, double* s_buf,
                                             unsigned* ctrl) {
  const int tid = threadIdx.x, bd = blockDim.x;
  const int par = epoch & 1u;
  void* me = P.mail[P.rank];
  double* s_part = s_buf + LL_BUF_VALS;
  __syncthreads();
  if (P.world > 1 && nv == 2 && P.world * vnblk * 2 <= LL_BUF_VALS) {
    const unsigned tag = ll_tag(epoch, id);
    if (tid < P.world * 2) {
      const int r = tid >> 1, i = tid & 1;
      ll_store(mb_x1(P.mail[r], par, P.rank, vblk) + 2 * i, s_in[i], tag);
    }
    const int per_rank = vnblk * 2, total = P.world * per_rank;
    for (int idx = tid; idx < total; idx += bd) {
      const int r = idx / per_rank, rem = idx - r * per_rank;
      const u64* slot = mb_x1(me, par, r, 0) + 2 * rem;
      double x = 0.0;
      KLERG_SPIN_UNTIL(ll_try_load(slot, tag, x), ctrl)
      s_buf[idx] = x;
    }
    __syncthreads();
    ll_reduce_staged(s_buf, P.world * vnblk, 2, max_mask, s_out, s_part);
    return;
  }
  {
    const unsigned tag = ll_tag(epoch, id);
    if (tid < nv) ll_store(mb_l1(me, par, vblk) + 2 * tid, s_in[tid], tag);
    const int total = vnblk * nv;
    for (int idx = tid; idx < total; idx += bd) {
      const int b = idx / nv, i = idx - b * nv;
      const u64* slot = mb_l1(me, par, b) + 2 * i;
      double x = 0.0;
      KLERG_SPIN_UNTIL(ll_try_load(slot, tag, x), ctrl)
      s_buf[idx] = x;
    }
    __syncthreads();
    ll_reduce_staged(s_buf, vnblk, nv, max_mask, s_out, s_part);
  }
  if (P.world > 1) {
    const unsigned tag = ll_tag(epoch, id + 0x20u);
    if (vblk == 0) {
      for (int t = tid; t < P.world * nv; t += bd) {
        const int r = t / nv, i = t - r * nv;
        ll_store(mb_r1(P.mail[r], par, P.rank) + 2 * i, s_out[i], tag);
      }
    }
    for (int idx = tid; idx < P.world * nv; idx += bd) {
      const int r = idx / nv, i = idx - r * nv;
      const u64* slot = mb_r1(me, par, r) + 2 * i;
      double x = 0.0;
      KLERG_SPIN_UNTIL(ll_try_load(slot, tag, x), ctrl)
      s_buf[idx] = x;
    }
    __syncthreads();
    if (tid < nv) {
      const bool is_max = (max_mask >> tid) & 1u;
      double v = is_max ? -INFINITY : 0.0;
      for (int r = 0; r < P.world; ++r) v = is_max ? fmax(v, s_buf[r * nv + tid]) : v + s_buf[r * nv + tid];
      s_out[tid] = v;
    }
    __syncthreads();
  }
}

// Exchange guard: slots of exchange number x are free once exchange x - 2 has been consumed, i.e. once the
// mailbox's xdone counter has reached x - 1 (signed distance: the counters wrap).
__device__ __forceinline__ void ll_wait_exchange_free(void* me, unsigned x, unsigned* ctrl) {
  const unsigned* xdone = mb_hdr(me) + 2;
  KLERG_SPIN_UNTIL((int)(ld_acquire_u32(xdone) - (x - 1u)) >= 0, ctrl)
}

}  // namespace klerg

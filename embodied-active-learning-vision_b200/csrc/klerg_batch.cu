// Robot.get_cost (klerg.py:686-710) for MANY candidate control sequences in ONE launch (BASELINE config 3: 1024
// candidates x horizon 50 x 1e6 samples).  Same arithmetic as eval_cost_kernel (klerg_fused.cuh) for every
// candidate; what changes is the organisation:
//   A  rollouts: candidate b is rolled out by CTA b % grid only (not by every CTA); its scaled post-step states go to
//      a small global table [B][H][D] that stays in L2.
//   B  forward pair pass.  The work is cut into units {group of 8 candidates} x {chunk of 4 samples per thread} and
//      CTA k takes the k-th equal share of the unit list (group-major): it touches at most a few groups, stages a
//      group's H*8 state rows once, and every thread runs the same number of full passes - no pass quantisation of a
//      per-CTA sample slice (1e6 / 148 = 6757 samples = 3.3 passes of 512 threads) and one block reduction per
//      touched group instead of one per group.  The workspace samples (1e6 x D floats) are read from L2.
//      q_base + q_iter goes to HBM (it is needed again once the normaliser is known: storing 4 B per sample and
//      candidate costs 0.6 ms of HBM time per GB, recomputing it a second MUFU-bound pair pass), per-CTA {sum, max}
//      per candidate to a partial table.
//   C  every candidate's {sum, max} over the CTAs that touched its group (candidate b by CTA b % grid, fixed order).
//   D  KL terms over the CTA's sample slice, one candidate per WARP at a time (p and log p of the slice staged in
//      shared memory; no block-wide barrier per group) -> partial table;   E  candidate b's cost by CTA b % grid.
// Four grid-wide meetings in total (flag-free tagged exchanges, klerg_ll.cuh) instead of two per group of 8.
#include "klerg_fused.cuh"

namespace klerg {

struct BatchArgs {
  int B;          // candidates
  float* xrows;   // [B][H][D] scaled post-step states
  float* bsum;    // [B] barrier sums
  double* part;   // [B][grid][2] {sum, max} of q_base + q_iter per CTA
  double* tot;    // [B][2]
  double* klp;    // [B][grid][2] KL terms per CTA
  int kl_cap;     // samples of p / log p staged in shared memory for phase D (0: read p from global memory)
};

template <int D>
__host__ __device__ inline SmemPlan plan_batch(int H, int S, int A, bool roll, int kl_cap) {
  constexpr int G = FUSED_MAXG;
  SmemPlan p{};
  size_t o = 0;
  p.u = o;    o = align16(o + sizeof(float) * (size_t)G * H * A);
  p.traj = o; o = align16(o + sizeof(float) * (size_t)G * (H + 1) * S);
  p.x2 = o;   o = align16(o + sizeof(u64) * (size_t)G * H * Row2<D>::DP);
  p.rx = o;   o = align16(o + sizeof(float) * (size_t)G * H * RowX<D>::NF);
  p.tile = o; o = align16(o + (roll ? sizeof(float) * rollout_rot_floats(G, H) : 0));
  p.red = o;  o = align16(o + sizeof(double) * (32 * 2 * FUSED_MAXG + 4 * FUSED_MAXG));
  p.misc = o; o = align16(o + 64 + sizeof(float) * (FUSED_MAXG + KLERG_MAX_D));
  p.ll = o;   o = align16(o + LL_SCRATCH_BYTES);
  p.xs = o;   o = align16(o + sizeof(float) * 2 * (size_t)kl_cap);
  p.total = o;
  return p;
}

template <int D>
__global__ void __launch_bounds__(512) eval_cost_batch_kernel(const __grid_constant__ EvalArgs a, const __grid_constant__ BatchArgs b) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int G = FUSED_MAXG, DP = Row2<D>::DP, NF = RowX<D>::NF;
  const int H = a.H, S = a.d.S, A = a.d.A, B = b.B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int vblk = blockIdx.x, vnblk = gridDim.x;
  const SmemPlan sp = plan_batch<D>(H, S, A, a.d.kind == KLERG_DYN_ROLL, b.kl_cap);
  float* s_u = (float*)(smem + sp.u);
  float* s_traj = (float*)(smem + sp.traj);
  u64* s_x2 = (u64*)(smem + sp.x2);
  float* s_rx = (float*)(smem + sp.rx);
  double* s_red = (double*)(smem + sp.red);
  double* s_world = s_red + 32 * 2 * FUSED_MAXG;
  double* s_in = s_world + 2 * FUSED_MAXG;
  unsigned* s_epoch = (unsigned*)(smem + sp.misc) + 1;
  float* s_bsum = (float*)(smem + sp.misc + 64);
  float* s_ctr = s_bsum + FUSED_MAXG;
  double* s_ll = (double*)(smem + sp.ll);
  unsigned* ctrl = ws_fused_ctrl(a.ws);
  void* me = a.peers.mail[a.peers.rank];

  pdl_wait_prior_grids();
  if (tid == 0) s_epoch[0] = ld_acquire_u32(mb_hdr(me));
  // common centre of the expanded pair form: the (scaled) initial state, from which every candidate's trajectory starts
  if (tid < D) s_ctr[tid] = a.x0[a.k.explr[tid]] * a.k.a[tid];
  __syncthreads();
  const unsigned epoch = s_epoch[0];
  int meeting = 0;
  // grid-wide meeting: an all-reduce of {nothing, a maximum}; consecutive meetings alternate slot parity
  auto meet = [&](double vmax) {
    if (tid == 0) {
      s_in[0] = 0.0;
      s_in[1] = vmax;
    }
    ll_allreduce(a.peers, vblk, vnblk, epoch + (unsigned)meeting, 0x10u + (unsigned)meeting, 2, 0x2u, s_in, s_world, s_ll, ctrl);
    ++meeting;
  };

  // ---- A: rollouts of the candidates b = vblk, vblk + grid, ... (up to G per pass through rollout_block) ----------
  float r2max = 0.f;  // largest squared distance of a state from the centre (decides the pair form for the whole launch)
  for (int b0 = vblk; b0 < B; b0 += vnblk * G) {
    int ng = 0;
    for (int g = 0; g < G && b0 + g * vnblk < B; ++g) ++ng;
    for (int e = tid; e < ng * H * A; e += blockDim.x) {
      const int g = e / (H * A);
      s_u[e] = a.u[(size_t)(b0 + g * vnblk) * H * A + (e - g * H * A)];
    }
    __syncthreads();
    rollout_block(a.d, a.bar, a.x0, a.R0, s_u, ng, H, s_traj, nullptr, nullptr, (float*)(smem + sp.tile), (float*)s_red, s_bsum,
                  nullptr);
    for (int e = tid; e < ng * H * D; e += blockDim.x) {
      const int g = e / (H * D), r = e - g * (H * D), t = r / D, d = r - t * D;
      const float x = s_traj[((size_t)g * (H + 1) + t + 1) * S + a.k.explr[d]] * a.k.a[d];
      b.xrows[(size_t)(b0 + g * vnblk) * H * D + r] = x;
      const float xc = x - s_ctr[d];
      r2max = fmaxf(r2max, xc * xc * (float)D);  // bound of |xc|^2 from the largest coordinate
    }
    if (tid < ng) b.bsum[b0 + tid * vnblk] = s_bsum[tid];
    __syncthreads();
  }
  __threadfence();
  {
    // one pair form for every candidate of the launch (a candidate's cost must not depend on its group mates)
    float m = r2max;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float* s_m = (float*)s_red;
    __syncthreads();
    if (lane == 0) s_m[warp] = m;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < nwarps; ++w) m = fmaxf(m, s_m[w]);
      s_m[0] = m;
    }
    __syncthreads();
    meet((double)s_m[0]);
  }
  const bool xform_all = s_world[1] <= (double)a.k.x_r2;

  // ---- B: forward pair pass over this CTA's share of the {group, sample chunk} units ----------------------------
  const bool narrow = a.N < (int64_t)4 * blockDim.x;                  // few samples: one packed pair per thread
  const int64_t CH = (int64_t)blockDim.x * (narrow ? 2 : 4);          // samples per chunk: one pass of the CTA
  const int64_t nch = (a.N + CH - 1) / CH;
  const int64_t U = nch * ((B + G - 1) / G);
  const int64_t u_lo = U * vblk / vnblk, u_hi = U * (vblk + 1) / vnblk;
  for (int64_t u = u_lo; u < u_hi;) {
    const int64_t grp = u / nch, c0 = u - grp * nch;
    const int64_t c1 = min(nch, c0 + (u_hi - u));
    const int g0 = (int)grp * G, ng = min(G, B - g0);
    for (int r = tid; r < ng * H; r += blockDim.x) {
      const float* row = b.xrows + ((size_t)g0 * H + r) * D;
      float x2n = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float x = __ldcg(row + d), xc = x - s_ctr[d];
        s_x2[r * DP + d] = pack2(x, x);
        s_rx[r * NF + d] = -2.f * xc;
        x2n = fmaf(xc, xc, x2n);
      }
#pragma unroll
      for (int d = D; d < DP; ++d) s_x2[r * DP + d] = pack2(0.f, 0.f);
      s_rx[r * NF + D] = x2n;
#pragma unroll
      for (int d = D + 1; d < NF; ++d) s_rx[r * NF + d] = 0.f;
    }
    StateRows sr;
    sr.x2 = s_x2; sr.rx = s_rx; sr.ctr = s_ctr;
    sr.xform = xform_all;
    __syncthreads();
    const int64_t lo = c0 * CH, hi = min(c1 * CH, a.ld);
    if (narrow)
      forward_candidates<D, 1>(a, sr, ng, H, lo, hi, s_red, s_in, a.v + (size_t)g0 * a.ld);
    else
      forward_candidates<D, 2>(a, sr, ng, H, lo, hi, s_red, s_in, a.v + (size_t)g0 * a.ld);
    __syncthreads();
    if (tid < 2 * ng) b.part[((size_t)(g0 + (tid >> 1)) * vnblk + vblk) * 2 + (tid & 1)] = s_in[tid];
    __syncthreads();
    u += c1 - c0;
  }
  __threadfence();
  meet(0.0);

  // ---- C: {sum, max} of every candidate over the CTAs that touched its group (fixed order), candidate by warp ----
  for (int c = vblk + vnblk * warp; c < B; c += vnblk * nwarps) {
    const int64_t g_lo = (int64_t)(c / G) * nch, g_hi = g_lo + nch;
    double sv = 0.0, mv = -INFINITY;
    for (int k = lane; k < vnblk; k += 32) {
      const int64_t k_lo = U * k / vnblk, k_hi = U * (k + 1) / vnblk;
      if (k_lo < g_hi && k_hi > g_lo && k_hi > k_lo) {
        sv += __ldcg(&b.part[((size_t)c * vnblk + k) * 2]);
        mv = fmax(mv, __ldcg(&b.part[((size_t)c * vnblk + k) * 2 + 1]));
      }
    }
    sv = warp_reduce(RED_SUM, sv);
    mv = warp_reduce(RED_MAX, mv);
    if (lane == 0) {
      b.tot[2 * c] = sv;
      b.tot[2 * c + 1] = mv;
      if (a.totals) {
        a.totals[2 * c] = sv;
        a.totals[2 * c + 1] = mv;
      }
    }
  }
  __threadfence();
  meet(0.0);

  // ---- D: KL terms over this CTA's sample slice, one candidate per warp at a time --------------------------------
  {
    int64_t lo, hi;
    cta_slice(a.N, a.ld, vblk, vnblk, lo, hi);
    const int n = (int)max((int64_t)0, min(hi, a.N) - lo);
    const bool staged = n <= b.kl_cap;
    float* s_p = (float*)(smem + sp.xs);
    float* s_lp = s_p + b.kl_cap;
    if (staged) {
      for (int i = tid; i < ((n + 3) & ~3); i += blockDim.x) {
        float pv = i < n ? __ldg(a.p + lo + i) : 1.f;
        if (pv != pv) pv = 1e-6f;
        s_p[i] = pv;
        s_lp[i] = __logf(pv);
      }
    }
    __syncthreads();
    for (int c = warp; c < B; c += nwarps) {
      const float vs = (float)__ldcg(&b.tot[2 * c]);
      const float rvs = 1.f / vs;
      const float maxc = fmaxf((float)__ldcg(&b.tot[2 * c + 1]) / vs, a.floor);
      const float* vrow = a.v + (size_t)c * a.ld + lo;
      double da = 0.0, dc = 0.0;
#pragma unroll 4
      for (int i0 = lane * 4; i0 < n; i0 += 128) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(vrow + i0));
        const float vv[4] = {t.x, t.y, t.z, t.w};
        float pv[4], lp[4];
        if (staged) {
          const float4 tp = *reinterpret_cast<const float4*>(s_p + i0), tl = *reinterpret_cast<const float4*>(s_lp + i0);
          pv[0] = tp.x; pv[1] = tp.y; pv[2] = tp.z; pv[3] = tp.w;
          lp[0] = tl.x; lp[1] = tl.y; lp[2] = tl.z; lp[3] = tl.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            pv[q] = (i0 + q < n) ? __ldg(a.p + lo + i0 + q) : 1.f;
            if (pv[q] != pv[q]) pv[q] = 1e-6f;
            lp[q] = __logf(pv[q]);
          }
        }
        float fa = 0.f, fc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (i0 + q < n) {
            float cq = fmaxf(vv[q] * rvs, a.floor);
            if (cq != cq) cq = 1e-6f * maxc;  // cost_norm: NaN in q -> 1e-6 (q = c / max c)
            fa = fmaf(pv[q], lp[q] - __logf(cq), fa);
            fc += cq;
          }
        }
        da += (double)fa;
        dc += (double)fc;
      }
      da = warp_reduce(RED_SUM, da);
      dc = warp_reduce(RED_SUM, dc);
      if (lane == 0) {
        b.klp[((size_t)c * vnblk + vblk) * 2] = da;
        b.klp[((size_t)c * vnblk + vblk) * 2 + 1] = dc;
      }
    }
  }
  __threadfence();
  meet(0.0);

  // ---- E: costs -----------------------------------------------------------------------------------------------------
  for (int c = vblk + vnblk * warp; c < B; c += vnblk * nwarps) {
    double s0 = 0.0, s1 = 0.0;
    for (int k = lane; k < vnblk; k += 32) {
      s0 += __ldcg(&b.klp[((size_t)c * vnblk + k) * 2]);
      s1 += __ldcg(&b.klp[((size_t)c * vnblk + k) * 2 + 1]);
    }
    s0 = warp_reduce(RED_SUM, s0);
    s1 = warp_reduce(RED_SUM, s1);
    if (lane == 0) {
      const double spv = a.p_stats[0];
      a.cost[c] = (float)(s0 / spv - log(spv) + log(s1)) + b.bsum[c];
    }
  }
  if (vblk == 0 && tid == 0) {
    st_release_u32(mb_hdr(me), epoch + 1u);
    if (a.fault_out) *a.fault_out = ctrl[5] ? 1.f : 0.f;
    __threadfence();
  }
  __syncthreads();
  pdl_launch_dependents();
}

template <int D>
static int launch_batch_d(EvalArgs& a, BatchArgs& b, cudaStream_t stream) {
  auto kernel = eval_cost_batch_kernel<D>;
  int64_t want = (a.N + 255) / 256;
  int nblk = want < sm_count() ? (int)want : sm_count();
  if (nblk > LL_MAXBLK) nblk = LL_MAXBLK;
  if (g_fused_opt.grid_limit > 0 && nblk > g_fused_opt.grid_limit) nblk = g_fused_opt.grid_limit;
  if (nblk < 1) nblk = 1;
  // 16 warps: four per scheduler (18 or 20 would leave schedulers unevenly loaded or spill registers)
  const int nthreads = 512;
  const int64_t per = ((a.N + nblk - 1) / nblk + 7) & ~(int64_t)7;  // cta_slice
  b.kl_cap = per <= 16384 ? (int)per : 0;
  const SmemPlan sp = plan_batch<D>(a.H, a.d.S, a.d.A, a.d.kind == KLERG_DYN_ROLL, b.kl_cap);
  if (sp.total > 200 * 1024) { set_error("eval_costs_batch: horizon too long for shared-memory staging"); return -1; }
  if (resident_ctas(kernel, nthreads, sp.total) < 1) { set_error("eval_costs_batch: kernel does not fit on an SM"); return -4; }
  return fused_launch(kernel, nblk, nthreads, sp.total, stream, "eval_cost_batch_kernel", false, a, b);
}

static size_t batch_scratch_layout(int64_t B, int64_t H, int D, size_t* off) {
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
  off[0] = take(sizeof(float) * (size_t)B * H * D);           // xrows
  off[1] = take(sizeof(float) * (size_t)B);                   // bsum
  off[2] = take(sizeof(double) * (size_t)B * LL_MAXBLK * 2);  // part
  off[3] = take(sizeof(double) * (size_t)B * 2);              // tot
  off[4] = take(sizeof(double) * (size_t)B * LL_MAXBLK * 2);  // klp
  return o;
}

}  // namespace klerg

using namespace klerg;

extern "C" size_t klerg_eval_costs_batch_scratch_bytes(int64_t B, int64_t H, int32_t D) {
  size_t off[5];
  return batch_scratch_layout(B, H, D, off);
}

extern "C" int klerg_eval_costs_batch(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                                      const float* x0, const float* R0, const float* u, int64_t B, int64_t H,
                                      const float* packed, int64_t N, int64_t ld, const float* q_base, const float* p,
                                      const double* p_stats, float floor, float* v_scratch, void* scratch,
                                      double* totals, float* cost, float* fault_out, void* workspace, void* stream) {
  EvalArgs a{};
  if (!make_kernel_dev(k, a.k) || !make_dyn(dyn, a.d) || !make_bar(bar, a.bar)) return -1;
  if (H < 1 || H > KLERG_MAX_H) { set_error("eval_costs_batch: H out of range"); return -1; }
  if (B < 1 || B > (1 << 20)) { set_error("eval_costs_batch: B out of range"); return -1; }
  if (N < 1 || ld < N || (ld & 3)) { set_error("eval_costs_batch: bad sample sizes"); return -1; }
  if (!workspace || !v_scratch || !scratch || !cost || !u || !x0 || !p || !p_stats) { set_error("eval_costs_batch: null argument"); return -1; }
  if (((uintptr_t)packed | (uintptr_t)p | (uintptr_t)v_scratch) & 15) { set_error("eval_costs_batch: packed, p and v_scratch must be 16-byte aligned"); return -1; }
  a.peers.world = 1; a.peers.rank = 0;
  for (int r = 0; r < MB_MAXW; ++r) a.peers.mail[r] = nullptr;
  a.peers.mail[0] = ws_fused_mailbox(workspace);
  a.independent = 0;
  a.x0 = x0; a.R0 = R0; a.u = u; a.G = FUSED_MAXG; a.H = (int)H; a.packed = packed; a.N = N; a.ld = ld; a.q_base = q_base;
  a.p = p; a.p_stats = p_stats; a.floor = floor; a.v = v_scratch; a.ws = workspace; a.totals = totals; a.cost = cost;
  a.K = 1; a.fault_out = fault_out;
  size_t off[5];
  batch_scratch_layout(B, H, a.k.D, off);
  BatchArgs b{};
  b.B = (int)B;
  b.xrows = (float*)((char*)scratch + off[0]);
  b.bsum = (float*)((char*)scratch + off[1]);
  b.part = (double*)((char*)scratch + off[2]);
  b.tot = (double*)((char*)scratch + off[3]);
  b.klp = (double*)((char*)scratch + off[4]);
  switch (a.k.D) {
    case 1: return launch_batch_d<1>(a, b, (cudaStream_t)stream);
    case 2: return launch_batch_d<2>(a, b, (cudaStream_t)stream);
    case 3: return launch_batch_d<3>(a, b, (cudaStream_t)stream);
    case 4: return launch_batch_d<4>(a, b, (cudaStream_t)stream);
    case 5: return launch_batch_d<5>(a, b, (cudaStream_t)stream);
    case 6: return launch_batch_d<6>(a, b, (cudaStream_t)stream);
    default: set_error("eval_costs_batch: D=%d not instantiated (1..6)", a.k.D); return -2;
  }
}

// Pairwise Gaussian-kernel passes of the KL-ergodic objective (sm_100a).
//
//   footprint_kernel : q_i = add_i + sum_j / max_j psi(state_j, s_i)     (a2, a3)
//   grad_kernel      : dgdx_t = sum_i w_i * (-(x_t - s_i)/scale) * psi    (a4)
//
// Both work in pre-scaled coordinates (x' = x*a_d, a_d = sqrt(0.5*log2e/|scale_d|))
// so one pair costs D FADD + D FFMA + 1 MUFU.EX2 (+1 FADD / +1 FMUL + D FFMA).
// States are staged in shared memory and read as warp-wide broadcasts; samples
// stream from the packed SoA array with coalesced 128-bit loads.
#include <cstdio>

#include "klerg_common.cuh"
#include "klerg_pair.cuh"

namespace klerg {

// ---------------------------------------------------------------------------
// pack: AoS [N][D] -> scaled SoA [D][ld]
// ---------------------------------------------------------------------------
__global__ void pack_samples_kernel(KernelDev k, const float* __restrict__ samples, int64_t N,
                                    float* __restrict__ packed, int64_t ld) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += stride) {
    for (int d = 0; d < k.D; ++d) {
      float v = (i < N) ? samples[i * k.D + d] * k.a[d] : 0.f;
      packed[(int64_t)d * ld + i] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// forward: footprint (sum) / spread (max) / both in one pass, packed f32x2 pair math (klerg_pair.cuh)
// ---------------------------------------------------------------------------
constexpr int FP_THREADS = 256;

// state rows staged per pass: 32 KB (D <= 4) / 36 KB (D <= 6): duplicated {x, x} rows of the difference form or,
// in the same buffer, the {-2 xc, |xc|^2} rows of the expanded form
template <int D>
struct FootChunk {
  static constexpr int ROWS = Row2<D>::DP <= 4 ? 1024 : 768;
  static_assert(sizeof(float) * RowX<D>::NF <= sizeof(u64) * Row2<D>::DP, "expanded rows must fit the staging buffer");
};

struct FootArgs {
  KernelDev k;
  const float* states;
  int64_t G, T, seg_stride;
  int64_t T_sum;  // MODE 2: rows [0, T_sum) enter the sum, all T rows the maximum
  const float* packed;
  int64_t N, ld;
  const float* add_in;
  float* out;
  int64_t out_stride;
  float* out_max;  // MODE 2: [G][out_stride] max_j psi
  double* totals;
  void* ws;
  int64_t ntiles;
  const int* gate;  // optional: the launch returns at once when *gate != 0 (the tensor-core pass did the work)
};

// Load SPT consecutive samples of one dimension starting at i0 as SPT/2 packed pairs.
template <int P>
__device__ __forceinline__ void load_sample_pairs(const float* __restrict__ p, int64_t i0, u64 (&s)[P]) {
  if constexpr (P == 2) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p + i0));
    s[0] = pack2(v.x, v.y);
    s[1] = pack2(v.z, v.w);
  } else {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p + i0));
    s[0] = pack2(v.x, v.y);
  }
}

// MODE 0: sum (traj_footprint_vec)  1: max (traj_spread_vec)  2: sum over the first T_sum rows and max over all
// rows in ONE pass over the squared distances (the history footprint q_base and the spread of get_target_dist
// visit the same memory-buffer rows: klerg.py:470-475 and :496).
//
// Every staged chunk of rows is evaluated in the expanded form around its middle row (klerg_pair.cuh: D + 2
// lane-ops per pair instead of 2D + 1); a chunk whose rows lie further than X_FORM_MAX_R2 from that row (in
// kernel widths) is evaluated in the exact difference form instead.
template <int D, int MODE, int P>
__global__ void __launch_bounds__(FP_THREADS) footprint_kernel(const FootArgs a) {
  constexpr int DP = Row2<D>::DP;
  constexpr int NF = RowX<D>::NF;
  constexpr int SPT = 2 * P;
  constexpr int TILE = FP_THREADS * SPT;
  constexpr int CHUNK = FootChunk<D>::ROWS;
  __shared__ __align__(16) u64 sh[CHUNK * DP];
  if (a.gate != nullptr && __ldg(a.gate) != 0) return;
  float* shx = reinterpret_cast<float*>(sh);
  const int tid = threadIdx.x;
  // chunk schedule: rows [0, T_sum) first, then [T_sum, T) (MODE 2; otherwise T_sum = T)
  const int64_t Tsum = MODE == 2 ? a.T_sum : a.T;
  const int nch1 = (int)((Tsum + CHUNK - 1) / CHUNK);
  const int nchunk = nch1 + (int)((a.T - Tsum + CHUNK - 1) / CHUNK);

  for (int64_t g = blockIdx.y; g < a.G; g += gridDim.y) {
    const float* st = a.states + g * a.seg_stride;
    double tsum = 0.0, tmax = -INFINITY;
    float ctr[D];  // centre of the staged chunk (scaled coordinates), same in every thread

    // stage chunk c; returns true if it was staged in the expanded form
    auto stage = [&](int c, int& rows) -> bool {
      const int64_t j0 = c < nch1 ? (int64_t)c * CHUNK : Tsum + (int64_t)(c - nch1) * CHUNK;
      const int64_t jend = c < nch1 ? Tsum : a.T;
      rows = (int)min((int64_t)CHUNK, jend - j0);
      const int64_t jm = j0 + rows / 2;
#pragma unroll
      for (int d = 0; d < D; ++d) ctr[d] = st[jm * a.k.S + a.k.explr[d]] * a.k.a[d];
      bool big = false;
      for (int j = tid; j < rows; j += FP_THREADS) {
        float r[NF];
        float x2n = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float xc = st[(j0 + j) * a.k.S + a.k.explr[d]] * a.k.a[d] - ctr[d];
          r[d] = -2.f * xc;
          x2n = fmaf(xc, xc, x2n);
        }
        r[D] = x2n;
#pragma unroll
        for (int d = D + 1; d < NF; ++d) r[d] = 0.f;
        big |= !(x2n <= a.k.x_r2);
        float4* dst = reinterpret_cast<float4*>(shx + (size_t)j * NF);
#pragma unroll
        for (int h = 0; h < NF / 4; ++h) dst[h] = make_float4(r[4 * h], r[4 * h + 1], r[4 * h + 2], r[4 * h + 3]);
      }
      if (!__syncthreads_or(big)) return true;
      for (int e = tid; e < rows * DP; e += FP_THREADS) {
        const int j = e / DP, d = e - j * DP;
        const float v = (d < D) ? st[(j0 + j) * a.k.S + a.k.explr[d]] * a.k.a[d] : 0.f;
        sh[e] = pack2(v, v);
      }
      __syncthreads();
      return false;
    };

    int rows_single = 0;
    bool x_single = false;
    if (nchunk == 1) {
      __syncthreads();
      x_single = stage(0, rows_single);
    }

    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const int64_t i0 = tile * TILE + (int64_t)tid * SPT;
      u64 s2[D][P], acc[P];
      float emin[SPT];
      const bool inb = i0 < a.ld;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (inb) {
          u64 t[P];
          load_sample_pairs<P>(a.packed + (int64_t)d * a.ld, i0, t);
#pragma unroll
          for (int q = 0; q < P; ++q) s2[d][q] = t[q];
        } else {
#pragma unroll
          for (int q = 0; q < P; ++q) s2[d][q] = pack2(0.f, 0.f);
        }
      }
#pragma unroll
      for (int q = 0; q < P; ++q) acc[q] = pack2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < SPT; ++q) emin[q] = INFINITY;

      // long state lists (history of 1e5 rows): every staged chunk is summed in a fresh accumulator and then
      // added to the running total, so the fp32 rounding error grows with sqrt(rows per chunk) + sqrt(chunks)
      // instead of sqrt(rows)
      u64 total[P];
#pragma unroll
      for (int q = 0; q < P; ++q) total[q] = pack2(0.f, 0.f);
      for (int c = 0; c < nchunk; ++c) {
        int rows = rows_single;
        bool xform = x_single;
        if (nchunk > 1) {
          __syncthreads();
          xform = stage(c, rows);
        }
        const bool summed = MODE != 2 || c < nch1;  // MODE 2: the rows behind T_sum only enter the maximum
        if (xform) {
          u64 sc[D][P], s2n[P];
#pragma unroll
          for (int q = 0; q < P; ++q) {
            const u64 c0 = pack2(ctr[0], ctr[0]);
            sc[0][q] = sub2(s2[0][q], c0);
            s2n[q] = mul2(sc[0][q], sc[0][q]);
#pragma unroll
            for (int d = 1; d < D; ++d) {
              sc[d][q] = sub2(s2[d][q], pack2(ctr[d], ctr[d]));
              s2n[q] = fma2(sc[d][q], sc[d][q], s2n[q]);
            }
          }
          if (summed)
            pair_forward_x<D, P, MODE>(shx, 0, rows, sc, s2n, acc, emin);
          else
            pair_forward_x<D, P, 1>(shx, 0, rows, sc, s2n, acc, emin);
        } else {
          if (summed)
            pair_forward<D, P, MODE>(sh, 0, rows, s2, acc, emin);
          else
            pair_forward<D, P, 1>(sh, 0, rows, s2, acc, emin);
        }
        if (MODE != 1 && nchunk > 1) {
#pragma unroll
          for (int q = 0; q < P; ++q) {
            total[q] = add2(total[q], acc[q]);
            acc[q] = pack2(0.f, 0.f);
          }
        }
      }
      if (MODE != 1 && nchunk > 1) {
#pragma unroll
        for (int q = 0; q < P; ++q) acc[q] = total[q];
      }

      // epilogue: scale, add base, store, local totals
      float o[SPT], om[SPT];
#pragma unroll
      for (int q = 0; q < P; ++q) unpack2(acc[q], o[2 * q], o[2 * q + 1]);
#pragma unroll
      for (int q = 0; q < SPT; ++q) {
        om[q] = ex2_neg(emin[q]) * a.k.inv_nu;  // T = 0: 2^-inf = 0
        float v = (MODE == 1) ? om[q] : o[q] * a.k.inv_nu;
        const int64_t i = i0 + q;
        if (a.add_in != nullptr && i < a.N) v += a.add_in[i];
        o[q] = v;
        if (i < a.N) {
          tsum += (double)v;
          tmax = fmax(tmax, (double)v);
        }
      }
      float* op = a.out + g * a.out_stride + i0;
      const bool vec = SPT == 4 && i0 + 3 < a.N && ((a.out_stride & 3) == 0);
      if (vec && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0)) {
        *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int q = 0; q < SPT; ++q)
          if (i0 + q < a.N) op[q] = o[q];
      }
      if (MODE == 2) {
        float* mp = a.out_max + g * a.out_stride + i0;
        if (vec && ((reinterpret_cast<uintptr_t>(a.out_max) & 15) == 0)) {
          *reinterpret_cast<float4*>(mp) = make_float4(om[0], om[1], om[2], om[3]);
        } else {
#pragma unroll
          for (int q = 0; q < SPT; ++q)
            if (i0 + q < a.N) mp[q] = om[q];
        }
      }
    }

    const int kinds[2] = {RED_SUM, RED_MAX};
    double vals[2] = {tsum, tmax};
    grid_reduce<2>(kinds, vals, ws_seg_partials(a.ws, g), ws_seg_counter(a.ws, g), blockIdx.x, gridDim.x,
                   a.totals + g * 2);
    __syncthreads();
  }
}

template <int D, int MODE>
static int launch_footprint_spt(const FootArgs& a0, cudaStream_t stream) {
  FootArgs a = a0;
  const int sms = sm_count();
  // choose samples-per-thread so that small workspaces still fill the chip
  int spt = 4;
  if (a.N * a.G < (int64_t)sms * FP_THREADS * 4 * 2) spt = 2;
  const int tile = FP_THREADS * spt;
  a.ntiles = (a.N + tile - 1) / tile;
  int64_t gx = a.ntiles < 1 ? 1 : a.ntiles;
  int64_t gy = a.G;
  if (gy > 65535) gy = 65535;
  // One or two tiles per CTA at the large sizes (1e7 samples = 9766 tiles): the hardware hands CTAs to SMs as they
  // free up, so an SM that is slower (it hosts the side-stream draw of the next step, engine.UniformPrefetch) simply
  // takes fewer, and the tail is one tile, not 1/8 of an SM's share (a static 8 CTAs per SM left ~8 % of the pass
  // to wave quantisation: 9766 tiles over 1184 CTAs are 8 or 9 tiles each).
  const int64_t cap = (int64_t)sms * 64 < MAXBLK ? (int64_t)sms * 64 : MAXBLK;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (spt == 4)
    footprint_kernel<D, MODE, 2><<<grid, FP_THREADS, 0, stream>>>(a);
  else
    footprint_kernel<D, MODE, 1><<<grid, FP_THREADS, 0, stream>>>(a);
  return check_launch("footprint_kernel");
}

template <int MODE>
static int launch_footprint_d(const FootArgs& a, cudaStream_t stream) {
  switch (a.k.D) {
    case 1: return launch_footprint_spt<1, MODE>(a, stream);
    case 2: return launch_footprint_spt<2, MODE>(a, stream);
    case 3: return launch_footprint_spt<3, MODE>(a, stream);
    case 4: return launch_footprint_spt<4, MODE>(a, stream);
    case 5: return launch_footprint_spt<5, MODE>(a, stream);
    case 6: return launch_footprint_spt<6, MODE>(a, stream);
    default: set_error("footprint: D=%d not instantiated (1..6)", a.k.D); return -2;
  }
}

// ---------------------------------------------------------------------------
// gradient
// ---------------------------------------------------------------------------
template <int SPT>
__device__ __forceinline__ void load_samples(const float* __restrict__ p, int64_t i0, float (&s)[SPT]) {
  static_assert(SPT == 4, "gradient tiles read four samples per lane");
  const float4 v = __ldg(reinterpret_cast<const float4*>(p + i0));
  s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
}

constexpr int GR_WARPS = 8;
constexpr int GR_THREADS = GR_WARPS * 32;
constexpr int GR_TILE = GR_THREADS * 4;  // samples per tile (w staged in smem)

struct GradArgs {
  KernelDev k;
  const float* states;  // [H][S]
  int64_t H;
  const float* packed;
  int64_t N, ld;
  const float* w;       // explicit importance ratio (FUSED = false)
  const float* v;       // FUSED: q_base + q_iter
  const float* p;       // FUSED
  const double* totals; // FUSED: [world][1][2]
  int world;
  float floor;
  double* grad_out;     // [H][D] (FUSED) or nullptr
  float* dgdx;          // [H][S] (explicit form) or nullptr
  double* kl_out;       // [2] FUSED
  void* ws;
  int64_t ntiles;
};

template <int D, int WT, bool FUSED>
__global__ void __launch_bounds__(GR_THREADS) grad_kernel(const GradArgs a) {
  __shared__ __align__(16) float sh_w[GR_TILE];
  __shared__ float sh_part[GR_WARPS * WT * D];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = (blockIdx.y * GR_WARPS + warp) * WT;

  float xs[WT][D];
#pragma unroll
  for (int kk = 0; kk < WT; ++kk) {
    const int64_t t = t0 + kk;
#pragma unroll
    for (int d = 0; d < D; ++d)
      xs[kk][d] = (t < a.H) ? a.states[t * a.k.S + a.k.explr[d]] * a.k.a[d] : 0.f;
  }
  float acc[WT][D];
#pragma unroll
  for (int kk = 0; kk < WT; ++kk)
#pragma unroll
    for (int d = 0; d < D; ++d) acc[kk][d] = 0.f;

  double vsum = 1.0, maxc = 1.0;
  if (FUSED) {
    double vmax;
    gather_totals(a.totals, a.world, 1, 0, vsum, vmax);
    maxc = fmax(vmax / vsum, (double)a.floor);
  }
  const float inv_vsum_f = (float)vsum;  // divide by the fp32 sum like the reference
  const float maxc_f = (float)maxc;
  double kl_a = 0.0, kl_c = 0.0;

  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    __syncthreads();
    {
      const int64_t i0 = tile * GR_TILE + (int64_t)tid * 4;
      float wv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t i = i0 + q;
        float wq = 0.f;
        if (i < a.N) {
          if (FUSED) {
            const float c = fmaxf(a.v[i] / inv_vsum_f, a.floor);
            const float pi = a.p[i];
            wq = pi * maxc_f / c;
            if (blockIdx.y == 0) {
              kl_a += (double)(pi * (logf(pi) - logf(c)));
              kl_c += (double)c;
            }
          } else {
            wq = a.w[i];
          }
        }
        wv[q] = wq;
      }
      *reinterpret_cast<float4*>(&sh_w[tid * 4]) = make_float4(wv[0], wv[1], wv[2], wv[3]);
    }
    __syncthreads();
#pragma unroll 2
    for (int it = 0; it < GR_TILE / 128; ++it) {
      const int off = (it * 32 + lane) * 4;
      const int64_t i0 = tile * GR_TILE + off;
      if (i0 >= a.ld) continue;
      float s[D][4];
#pragma unroll
      for (int d = 0; d < D; ++d) load_samples<4>(a.packed + (int64_t)d * a.ld, i0, s[d]);
      const float4 w4 = *reinterpret_cast<const float4*>(&sh_w[off]);
      const float wq[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int kk = 0; kk < WT; ++kk) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float df[D];
          df[0] = xs[kk][0] - s[0][q];
          float e = -df[0] * df[0];
#pragma unroll
          for (int d = 1; d < D; ++d) {
            df[d] = xs[kk][d] - s[d][q];
            e = fmaf(-df[d], df[d], e);
          }
          const float wp = wq[q] * ex2_approx(e);
#pragma unroll
          for (int d = 0; d < D; ++d) acc[kk][d] = fmaf(wp, df[d], acc[kk][d]);
        }
      }
    }
  }

  // lanes -> warp totals -> block partial [H_blk][D] in global (doubles)
#pragma unroll
  for (int kk = 0; kk < WT; ++kk)
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float v = warp_sum_f(acc[kk][d]);
      if (lane == 0) sh_part[(warp * WT + kk) * D + d] = v;
    }
  __syncthreads();
  // this block covers states [blockIdx.y*GR_WARPS*WT, +GR_WARPS*WT)
  const int nval = GR_WARPS * WT * D;
  const int nblkx = gridDim.x;
  double* parts = ws_grad_partials(a.ws);  // [blockIdx.x][H*D]
  const int64_t HD = a.H * D;
  for (int e = tid; e < nval; e += GR_THREADS) {
    const int64_t t = (int64_t)blockIdx.y * GR_WARPS * WT + e / D;
    if (t < a.H) parts[(size_t)blockIdx.x * HD + t * D + (e % D)] = (double)sh_part[e];
  }
  // KL partials ride along in the misc partial area (only blockIdx.y == 0 contributes)
  const int kinds[2] = {RED_SUM, RED_SUM};
  double vals[2] = {kl_a, kl_c};
  double kl_tmp[2];
  __shared__ double sh_kl[2];
  const int blk_lin = blockIdx.y * gridDim.x + blockIdx.x;
  const int nblk = gridDim.x * gridDim.y;
  const bool last = grid_reduce<2>(kinds, vals, ws_misc_partials(a.ws), ws_misc_counter(a.ws, 0), blk_lin, nblk,
                                   sh_kl);
  (void)kl_tmp;
  if (last) {
    __syncthreads();
    // fixed-order sum over x-blocks of the gradient partials
    for (int64_t e = tid; e < HD; e += GR_THREADS) {
      double sacc = 0.0;
      for (int b = 0; b < nblkx; ++b) sacc += __ldcg(&parts[(size_t)b * HD + e]);
      const int d = (int)(e % D);
      const double gval = sacc * (double)a.k.gfac[d];
      if (a.grad_out) a.grad_out[e] = gval;
      if (a.dgdx) a.dgdx[(e / D) * a.k.S + a.k.explr[d]] = (float)gval;
    }
    if (FUSED && tid == 0 && a.kl_out) {
      a.kl_out[0] = sh_kl[0];
      a.kl_out[1] = sh_kl[1];
    }
  }
}

template <int D, int WT, bool FUSED>
static int launch_grad(GradArgs a, cudaStream_t stream) {
  a.ntiles = (a.N + GR_TILE - 1) / GR_TILE;
  const int gy = (int)((a.H + GR_WARPS * WT - 1) / (GR_WARPS * WT));
  int64_t gx = a.ntiles < 1 ? 1 : a.ntiles;
  int64_t cap = (int64_t)sm_count() * 2 / gy;
  if (cap < 1) cap = 1;
  if (cap > GRAD_MAXBLK) cap = GRAD_MAXBLK;
  if (gx > cap) gx = cap;
  if ((int64_t)gx * gy > MAXBLK) { set_error("grad: grid too large"); return -2; }
  grad_kernel<D, WT, FUSED><<<dim3((unsigned)gx, (unsigned)gy), GR_THREADS, 0, stream>>>(a);
  return check_launch("grad_kernel");
}

template <int D, bool FUSED>
static int launch_grad_wt(const GradArgs& a, cudaStream_t stream) {
  const int64_t per_warp = (a.H + GR_WARPS - 1) / GR_WARPS;
  if (per_warp <= 1) return launch_grad<D, 1, FUSED>(a, stream);
  if (per_warp <= 2) return launch_grad<D, 2, FUSED>(a, stream);
  if (per_warp <= 4) return launch_grad<D, 4, FUSED>(a, stream);
  return launch_grad<D, 7, FUSED>(a, stream);
}

template <bool FUSED>
static int launch_grad_d(const GradArgs& a, cudaStream_t stream) {
  switch (a.k.D) {
    case 1: return launch_grad_wt<1, FUSED>(a, stream);
    case 2: return launch_grad_wt<2, FUSED>(a, stream);
    case 3: return launch_grad_wt<3, FUSED>(a, stream);
    case 4: return launch_grad_wt<4, FUSED>(a, stream);
    case 5: return launch_grad_wt<5, FUSED>(a, stream);
    case 6: return launch_grad_wt<6, FUSED>(a, stream);
    default: set_error("gradient: D=%d not instantiated (1..6)", a.k.D); return -2;
  }
}

// ---------------------------------------------------------------------------
// a1: the kernel matrix itself (psi_fn, klerg_utils.py:7-10) and its derivative for one state
// (dpsi_dx_fn, :12-15).  Not on the planner's path (it never materialises N x T); exported for the
// callers of the reference's free functions.
// ---------------------------------------------------------------------------
__global__ void psi_matrix_kernel(KernelDev k, const float* __restrict__ states, int64_t T,
                                  const float* __restrict__ samples, int64_t N, float* __restrict__ psi,
                                  float* __restrict__ dpsi) {
  const int64_t total = N * T;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / T, j = e - i * T;
    float acc = 0.f, df[KLERG_MAX_D];
    for (int d = 0; d < k.D; ++d) {
      df[d] = (states[j * k.S + k.explr[d]] - samples[i * k.D + d]) * k.a[d];
      acc = fmaf(df[d], df[d], acc);
    }
    const float v = ex2_approx(-acc) * k.inv_nu;
    if (psi) psi[e] = v;
    if (dpsi)  // -(x - s)/|scale| * psi, one row of D values per (sample, state)
      for (int d = 0; d < k.D; ++d) dpsi[e * k.D + d] = df[d] * k.gfac[d] * (v / k.inv_nu);
  }
}

}  // namespace klerg

using namespace klerg;

extern "C" int klerg_psi_matrix(const klerg_kernel_spec* k, const float* states, int64_t T, const float* samples,
                                int64_t N, float* psi, float* dpsi, void* stream) {
  KernelDev kd;
  if (!make_kernel_dev(k, kd)) return -1;
  if (T < 0 || N < 0) { set_error("psi_matrix: bad sizes"); return -1; }
  if (T * N == 0) return 0;
  int64_t blocks = (T * N + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  psi_matrix_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(kd, states, T, samples, N, psi, dpsi);
  return check_launch("psi_matrix_kernel");
}

extern "C" int klerg_pack_samples(const klerg_kernel_spec* k, const float* samples, int64_t N, float* packed,
                                  int64_t ld, void* stream) {
  KernelDev kd;
  if (!make_kernel_dev(k, kd)) return -1;
  if (ld < N || (ld & 3)) { set_error("pack_samples: ld must be >= N and a multiple of 4"); return -1; }
  if (ld == 0) return 0;
  int64_t blocks = (ld + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  pack_samples_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(kd, samples, N, packed, ld);
  return check_launch("pack_samples_kernel");
}

extern "C" int klerg_footprint(const klerg_kernel_spec* k, int mode, const float* states, int64_t G, int64_t T,
                               int64_t seg_stride, const float* packed, int64_t N, int64_t ld,
                               const float* add_in, float* out, int64_t out_stride, double* totals,
                               void* workspace, void* stream) {
  KernelDev kd;
  if (!make_kernel_dev(k, kd)) return -1;
  if (G < 1 || N < 0 || T < 0 || ld < N || (ld & 3)) { set_error("footprint: bad sizes"); return -1; }
  if (!workspace || !totals || !out) { set_error("footprint: null output/workspace"); return -1; }
  FootArgs a{kd, states, G, T, seg_stride, T, packed, N, ld, add_in, out, out_stride, nullptr, totals, workspace, 0};
  if (mode == 0) return launch_footprint_d<0>(a, (cudaStream_t)stream);
  if (mode == 1) return launch_footprint_d<1>(a, (cudaStream_t)stream);
  set_error("footprint: mode must be 0 (sum) or 1 (max)");
  return -1;
}

extern "C" int klerg_footprint_sum_max(const klerg_kernel_spec* k, const float* states, int64_t T, int64_t T_sum,
                                       const float* packed, int64_t N, int64_t ld, float* out_sum, float* out_max,
                                       double* totals, void* workspace, void* stream) {
  KernelDev kd;
  if (!make_kernel_dev(k, kd)) return -1;
  if (N < 0 || T < 0 || T_sum < 0 || T_sum > T || ld < N || (ld & 3)) { set_error("footprint_sum_max: bad sizes"); return -1; }
  if (!workspace || !totals || !out_sum || !out_max) { set_error("footprint_sum_max: null output/workspace"); return -1; }
  FootArgs a{kd, states, 1, T, T * kd.S, T_sum, packed, N, ld, nullptr, out_sum, ld, out_max, totals, workspace, 0};
  return launch_footprint_d<2>(a, (cudaStream_t)stream);
}

namespace klerg {
int launch_footprint_sum_max_gated(const KernelDev& kd, const float* states, int64_t T, int64_t T_sum, const float* packed,
                                   int64_t N, int64_t ld, float* out_sum, float* out_max, double* totals, void* workspace,
                                   const int* gate, cudaStream_t stream) {
  FootArgs a{kd, states, 1, T, T * kd.S, T_sum, packed, N, ld, nullptr, out_sum, ld, out_max, totals, workspace, 0, gate};
  return launch_footprint_d<2>(a, stream);
}
}  // namespace klerg

extern "C" int klerg_kl_gradient(const klerg_kernel_spec* k, const float* states, int64_t H, const float* packed,
                                 int64_t N, int64_t ld, const float* w, float* dgdx, void* workspace,
                                 void* stream) {
  KernelDev kd;
  if (!make_kernel_dev(k, kd)) return -1;
  if (H < 1 || H > KLERG_MAX_H) { set_error("kl_gradient: H out of range"); return -1; }
  cudaError_t e = cudaMemsetAsync(dgdx, 0, sizeof(float) * H * kd.S, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("kl_gradient: memset failed: %s", cudaGetErrorString(e)); return -4; }
  GradArgs a{};
  a.k = kd; a.states = states; a.H = H; a.packed = packed; a.N = N; a.ld = ld; a.w = w;
  a.dgdx = dgdx; a.ws = workspace; a.world = 1; a.floor = 0.f;
  return launch_grad_d<false>(a, (cudaStream_t)stream);
}

extern "C" int klerg_kl_gradient_fused(const klerg_kernel_spec* k, const float* states, int64_t H,
                                       const float* packed, int64_t N, int64_t ld, const float* v,
                                       const double* totals, int world, const float* p, float floor,
                                       double* grad_part, double* kl_part, void* workspace, void* stream) {
  KernelDev kd;
  if (!make_kernel_dev(k, kd)) return -1;
  if (H < 1 || H > KLERG_MAX_H) { set_error("kl_gradient_fused: H out of range"); return -1; }
  GradArgs a{};
  a.k = kd; a.states = states; a.H = H; a.packed = packed; a.N = N; a.ld = ld; a.v = v; a.p = p;
  a.totals = totals; a.world = world; a.floor = floor; a.grad_out = grad_part; a.kl_out = kl_part;
  a.ws = workspace;
  return launch_grad_d<true>(a, (cudaStream_t)stream);
}

// Fingerprint belief update on the device (sm_100a): FingerprintDist.update_prior of the reference
// (franka_test/scripts/dist_modules/fingerprint_module.py:539-589) with meas_footprint_vec (:417-424) and the numpy
// renormalize it imports (control/klerg_utils.py:41-54).  The producer of the K belief targets p_k of BASELINE config 5.
//
// For n new measurements (location loc_j, processed value val_j) and a belief {prior, prior_var} over a 50^D grid:
//   m[g][j]   = exp(-0.5 sum_d (loc_jd - grid_gd)^2 / |std|),  std = max(scale / 2, 1e-6)      the same pair kernel
//   mm[g][j]  = renormalize over g (per measurement):  max(m / S_j, 1e-6) / max_g(...)
//   mv[g]     = renormalize(mean_j mm[g][j]),  meas_var = mv * (scale - 50 scale) + 50 scale
//   post_var  = 1 / (1 / prior_var + n / meas_var),  post = post_var * (prior / prior_var + sum_j (val_j/2 + 0.5) / meas_var)
// The reference computes this in numpy float64; so does this file (double precision: the grids are 2.5e3 .. 1.25e5
// points and a handful of measurements, nothing here is throughput-bound).  Reductions are fixed-order.
#include "klerg_common.cuh"

namespace klerg {

constexpr int BF_THREADS = 256;
constexpr int BF_MAXBLK = 256;

struct BeliefArgs {
  const double* grid;   // [G][D]
  int64_t G;
  int D, n;
  const double* locs;   // [n][D]
  double inv_std;       // 1 / |std|
  double scale, meas_sum;
  const double* prior;
  const double* prior_var;
  double* post;
  double* post_var;
  double* col_part;     // [n][BF_MAXBLK][2]
  double* col;          // [n][2] {S_j, max_g max(m / S_j, 1e-6)}
  double* mv;           // [G]
  double* mv_part;      // [BF_MAXBLK][2]
  double* mv_tot;       // [2]
};

__device__ __forceinline__ double belief_m(const BeliefArgs& a, int64_t g, int j) {
  double acc = 0.0;
  for (int d = 0; d < a.D; ++d) {
    const double df = a.locs[(size_t)j * a.D + d] - a.grid[(size_t)g * a.D + d];
    acc += df * df * a.inv_std;
  }
  return exp(-0.5 * acc);
}

// block reduction of {sum, max}; result valid in thread 0
__device__ __forceinline__ void block_sum_max(double& s, double& m) {
  __shared__ double sh[2][BF_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s = warp_reduce(RED_SUM, s);
  m = warp_reduce(RED_MAX, m);
  __syncthreads();
  if (lane == 0) {
    sh[0][warp] = s;
    sh[1][warp] = m;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BF_THREADS / 32; ++w) {
      s += sh[0][w];
      m = fmax(m, sh[1][w]);
    }
  }
}

// per measurement j (blockIdx.y) and grid chunk (blockIdx.x): {sum_g m, max_g m}
__global__ void __launch_bounds__(BF_THREADS) belief_col_partial(const BeliefArgs a) {
  const int j = blockIdx.y;
  double s = 0.0, m = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * BF_THREADS + threadIdx.x; g < a.G; g += (int64_t)gridDim.x * BF_THREADS) {
    const double v = belief_m(a, g, j);
    s += v;
    m = fmax(m, v);
  }
  block_sum_max(s, m);
  if (threadIdx.x == 0) {
    a.col_part[((size_t)j * BF_MAXBLK + blockIdx.x) * 2] = s;
    a.col_part[((size_t)j * BF_MAXBLK + blockIdx.x) * 2 + 1] = m;
  }
}
__global__ void belief_col_final(const BeliefArgs a, int nblk) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  double s = 0.0, m = 0.0;
  for (int b = 0; b < nblk; ++b) {
    s += a.col_part[((size_t)j * BF_MAXBLK + b) * 2];
    m = fmax(m, a.col_part[((size_t)j * BF_MAXBLK + b) * 2 + 1]);
  }
  a.col[2 * j] = s;
  a.col[2 * j + 1] = fmax(m / s, 1e-6);  // the largest clipped entry of the column
}
// mv[g] = mean_j renormalised column entry; partial {sum, max} over g
__global__ void __launch_bounds__(BF_THREADS) belief_mean(const BeliefArgs a) {
  double s = 0.0, m = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * BF_THREADS + threadIdx.x; g < a.G; g += (int64_t)gridDim.x * BF_THREADS) {
    double acc = 0.0;
    for (int j = 0; j < a.n; ++j) {
      const double c = fmax(belief_m(a, g, j) / a.col[2 * j], 1e-6);
      acc += exp(log(c) - log(a.col[2 * j + 1]));  // exp(log x - max log x), as renormalize() forms it
    }
    const double v = acc / (double)a.n;
    a.mv[g] = v;
    s += v;
    m = fmax(m, v);
  }
  block_sum_max(s, m);
  if (threadIdx.x == 0) {
    a.mv_part[blockIdx.x * 2] = s;
    a.mv_part[blockIdx.x * 2 + 1] = m;
  }
}
__global__ void belief_mean_final(const BeliefArgs a, int nblk) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0, m = 0.0;
  for (int b = 0; b < nblk; ++b) {
    s += a.mv_part[2 * b];
    m = fmax(m, a.mv_part[2 * b + 1]);
  }
  a.mv_tot[0] = s;
  a.mv_tot[1] = fmax(m / s, 1e-6);
}
__global__ void __launch_bounds__(BF_THREADS) belief_posterior(const BeliefArgs a) {
  const double S = a.mv_tot[0], lmax = log(a.mv_tot[1]);
  for (int64_t g = (int64_t)blockIdx.x * BF_THREADS + threadIdx.x; g < a.G; g += (int64_t)gridDim.x * BF_THREADS) {
    const double r = exp(log(fmax(a.mv[g] / S, 1e-6)) - lmax);
    const double meas_var = r * (a.scale - 50.0 * a.scale) + 50.0 * a.scale;
    const double pv = a.prior_var[g];
    const double post_var = 1.0 / (1.0 / pv + (double)a.n / meas_var);
    a.post_var[g] = post_var;
    a.post[g] = post_var * (a.prior[g] / pv + a.meas_sum / meas_var);
  }
}

}  // namespace klerg

using namespace klerg;

extern "C" size_t klerg_belief_scratch_bytes(int64_t G, int32_t n) {
  return sizeof(double) * ((size_t)n * BF_MAXBLK * 2 + (size_t)n * 2 + (size_t)G + BF_MAXBLK * 2 + 2);
}

extern "C" int klerg_belief_update(const double* grid, int64_t G, int32_t D, const double* locs, int32_t n, double scale,
                                   double meas_sum, const double* prior, const double* prior_var, double* posterior,
                                   double* posterior_var, void* scratch, void* stream) {
  if (!grid || !locs || !prior || !prior_var || !posterior || !posterior_var || !scratch) { set_error("belief_update: null argument"); return -1; }
  if (G < 1 || D < 1 || D > KLERG_MAX_D || n < 1 || n > 65535) { set_error("belief_update: bad sizes (G=%lld D=%d n=%d)", (long long)G, D, n); return -1; }
  if (!(scale > 0.0)) { set_error("belief_update: scale must be positive"); return -1; }
  BeliefArgs a{};
  a.grid = grid; a.G = G; a.D = D; a.n = n; a.locs = locs;
  const double std_ = scale / 2.0 < 1e-6 ? 1e-6 : scale / 2.0;
  a.inv_std = 1.0 / std_;
  a.scale = scale; a.meas_sum = meas_sum; a.prior = prior; a.prior_var = prior_var; a.post = posterior; a.post_var = posterior_var;
  double* s = (double*)scratch;
  a.col_part = s; s += (size_t)n * BF_MAXBLK * 2;
  a.col = s; s += (size_t)n * 2;
  a.mv = s; s += (size_t)G;
  a.mv_part = s; s += BF_MAXBLK * 2;
  a.mv_tot = s;
  int nblk = (int)((G + BF_THREADS - 1) / BF_THREADS);
  if (nblk > BF_MAXBLK) nblk = BF_MAXBLK;
  cudaStream_t st = (cudaStream_t)stream;
  belief_col_partial<<<dim3((unsigned)nblk, (unsigned)n), BF_THREADS, 0, st>>>(a);
  if (int e = check_launch("belief_col_partial")) return e;
  belief_col_final<<<(n + 127) / 128, 128, 0, st>>>(a, nblk);
  if (int e = check_launch("belief_col_final")) return e;
  belief_mean<<<nblk, BF_THREADS, 0, st>>>(a);
  if (int e = check_launch("belief_mean")) return e;
  belief_mean_final<<<1, 32, 0, st>>>(a, nblk);
  if (int e = check_launch("belief_mean_final")) return e;
  belief_posterior<<<nblk, BF_THREADS, 0, st>>>(a);
  return check_launch("belief_posterior");
}

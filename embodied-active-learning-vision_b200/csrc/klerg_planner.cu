// Planner-side kernels: vector statistics / renormalise / cost_norm, KL cost,
// target-density weighting, rollout + barrier + linearisation, adjoint sweep,
// memory-buffer row gather.  All tiny next to the pairwise passes; written so a
// whole planner iteration stays on the device (no host arithmetic).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "klerg_common.cuh"
#include "klerg_dyn.cuh"

namespace klerg {

// ---------------------------------------------------------------------------
// host utilities
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static long long g_launches = 0;

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -4;
  }
  return 0;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached = n;
  }
  return cached;
}

int g_exact_pairs = 0;
int g_saturate_milli = 0;

bool make_kernel_dev(const klerg_kernel_spec* k, KernelDev& out) {
  if (!k) { set_error("kernel spec is null"); return false; }
  if (k->D < 1 || k->D > KLERG_MAX_D || k->S < 1) { set_error("kernel spec: D=%d S=%d out of range", k->D, k->S); return false; }
  out.D = k->D;
  out.S = k->S;
  const double nu = (double)k->nu;
  out.inv_nu = (float)(1.0 / nu);
  out.x_r2 = g_exact_pairs ? -1.f : 100.f;  // radius^2 of the expanded pair form (X_FORM_MAX_R2, klerg_pair.cuh)
  for (int d = 0; d < KLERG_MAX_D; ++d) {
    out.explr[d] = 0;
    out.a[d] = 0.f;
    out.gfac[d] = 0.f;
  }
  for (int d = 0; d < k->D; ++d) {
    if (k->explr[d] < 0 || k->explr[d] >= k->S) { set_error("kernel spec: explr[%d]=%d outside [0,%d)", d, k->explr[d], k->S); return false; }
    const double sc = fabs((double)k->scale[d]);
    if (!(sc > 0.0)) { set_error("kernel spec: scale[%d] must be non-zero", d); return false; }
    const double a = sqrt(HALF_LOG2E / sc);
    out.explr[d] = k->explr[d];
    out.a[d] = (float)a;
    out.gfac[d] = (float)(-1.0 / ((double)(float)a * sc * nu));
  }
  return true;
}

// ---------------------------------------------------------------------------
// vector statistics, renormalize, cost_norm
// ---------------------------------------------------------------------------
constexpr int EW_THREADS = 256;

static int ew_blocks(int64_t N) {
  int64_t b = (N + EW_THREADS * 4 - 1) / (EW_THREADS * 4);
  const int64_t cap = (int64_t)sm_count() * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__global__ void __launch_bounds__(EW_THREADS) vector_stats_kernel(const float* __restrict__ x, int64_t N,
                                                                    double* stats, void* ws) {
  double s = 0.0, mx = -INFINITY, mn = INFINITY, nn = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (v != v) { nn += 1.0; continue; }
    s += (double)v;
    mx = fmax(mx, (double)v);
    mn = fmin(mn, (double)v);
  }
  const int kinds[4] = {RED_SUM, RED_MAX, RED_MIN, RED_SUM};
  double vals[4] = {s, mx, mn, nn};
  grid_reduce<4>(kinds, vals, ws_misc_partials(ws), ws_misc_counter(ws, 1), blockIdx.x, gridDim.x, stats);
}

// out = c / max c,  c = max(x / sum, floor); stats = {sum, max, ...} of x
__global__ void renorm_apply_kernel(const float* __restrict__ x, int64_t N, const double* __restrict__ stats,
                                    float floor, float* __restrict__ out) {
  const float sum = (float)stats[0];
  const float maxc = fmaxf((float)stats[1] / sum, floor);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = fmaxf(x[i] / sum, floor) / maxc;
}

// NaN -> 1e-6 and accumulate the sum of the sanitised vector
__global__ void __launch_bounds__(EW_THREADS) cost_norm_pass1(float* __restrict__ x, int64_t N, double* stats, void* ws) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    if (v != v) { v = 1e-6f; x[i] = v; }
    s += (double)v;
  }
  const int kinds[1] = {RED_SUM};
  double vals[1] = {s};
  grid_reduce<1>(kinds, vals, ws_misc_partials(ws), ws_misc_counter(ws, 1), blockIdx.x, gridDim.x, stats);
}

__global__ void cost_norm_pass2(float* __restrict__ x, int64_t N, const double* __restrict__ stats) {
  const float sum = (float)stats[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = x[i] / sum;
}

// ---------------------------------------------------------------------------
// KL cost of get_cost (klerg.py:694-699) for G candidates
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) kl_cost_partial_kernel(const float* __restrict__ v, int64_t v_stride,
                                                                       int64_t G, int64_t N,
                                                                       const double* __restrict__ totals, int world,
                                                                       const float* __restrict__ p, float floor,
                                                                       double* kl_part, void* ws) {
  for (int64_t g = blockIdx.y; g < G; g += gridDim.y) {
    double vsum, vmax;
    gather_totals(totals, world, G, g, vsum, vmax);
    const float vs = (float)vsum;
    const float maxc = fmaxf((float)vmax / vs, floor);
    const float* vg = v + g * v_stride;
    double sa = 0.0, sc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
      float c = fmaxf(vg[i] / vs, floor);
      if (c != c) c = 1e-6f * maxc;  // cost_norm: NaN in q -> 1e-6 (q = c/maxc)
      float pi = p[i];
      if (pi != pi) pi = 1e-6f;
      sa += (double)(pi * (logf(pi) - logf(c)));
      sc += (double)c;
    }
    const int kinds[2] = {RED_SUM, RED_SUM};
    double vals[2] = {sa, sc};
    grid_reduce<2>(kinds, vals, ws_seg_partials(ws, g), ws_seg_counter(ws, g), blockIdx.x, gridDim.x, kl_part + g * 2);
    __syncthreads();
  }
}

__global__ void kl_cost_final_kernel(const double* __restrict__ kl_part, int world, int64_t G,
                                     const double* __restrict__ p_stats, const float* __restrict__ barrier_sum,
                                     float* __restrict__ cost) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double sa = 0.0, sc = 0.0;
  for (int r = 0; r < world; ++r) {
    sa += kl_part[((size_t)r * G + g) * 2 + 0];
    sc += kl_part[((size_t)r * G + g) * 2 + 1];
  }
  const double sp = p_stats[0];
  const double dkl = sa / sp - log(sp) + log(sc);
  cost[g] = (float)dkl + (barrier_sum ? barrier_sum[g] : 0.f);
}

// ---------------------------------------------------------------------------
// target-density weighting (klerg.py:452-486)
// ---------------------------------------------------------------------------
struct LimDev {
  int D;
  float lo[KLERG_MAX_D], hi[KLERG_MAX_D];
};

__device__ __forceinline__ bool outside_box(const float* __restrict__ samples, int64_t i, const LimDev& L) {
  bool out = false;
  for (int d = 0; d < L.D; ++d) {
    const float s = samples[i * L.D + d];
    out |= (s < L.lo[d]) | (s > L.hi[d]);
  }
  return out;
}

__global__ void __launch_bounds__(EW_THREADS) target_stage1_kernel(const float* __restrict__ samples, LimDev L, int64_t N,
                                                                     const float* __restrict__ spread,
                                                                     const float* __restrict__ p, double* acc, void* ws) {
  double smax = -INFINITY, sin_ = 0.0, nout = 0.0, pmin = INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    pmin = fmin(pmin, (double)p[i]);
    if (spread) {
      const float s = spread[i];
      smax = fmax(smax, (double)s);
      if (outside_box(samples, i, L)) nout += 1.0;
      else sin_ += (double)s;
    }
  }
  if (!spread) smax = 1.0;
  const int kinds[4] = {RED_MAX, RED_SUM, RED_SUM, RED_MIN};
  double vals[4] = {smax, sin_, nout, pmin};
  grid_reduce<4>(kinds, vals, ws_misc_partials(ws), ws_misc_counter(ws, 1), blockIdx.x, gridDim.x, acc);
}

__global__ void __launch_bounds__(EW_THREADS) target_stage2_kernel(int mode, const float* __restrict__ samples, LimDev L,
                                                                     int64_t N, const float* __restrict__ spread,
                                                                     const float* __restrict__ p,
                                                                     const double* __restrict__ expo,
                                                                     float* __restrict__ p2, double* acc, void* ws) {
  const float ex = (float)expo[0];
  const float pmin = (float)expo[1];
  const float smax = (float)expo[2];
  double s = 0.0, mx = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float v = p[i];
    if (mode == 0) {
      v = powf(v, ex);
    } else if (mode == 1) {
      float sp = 0.f;  // empty buffer: spread = zeros(1)
      if (spread) sp = outside_box(samples, i, L) ? 1.f : spread[i] / smax;
      v = v + (1.f - sp) * pmin;
    }
    p2[i] = v;
    s += (double)v;
    mx = fmax(mx, (double)v);
  }
  const int kinds[2] = {RED_SUM, RED_MAX};
  double vals[2] = {s, mx};
  grid_reduce<2>(kinds, vals, ws_misc_partials(ws), ws_misc_counter(ws, 1), blockIdx.x, gridDim.x, acc);
}

__global__ void __launch_bounds__(EW_THREADS) target_stage3_kernel(const float* __restrict__ p2, int64_t N,
                                                                     const double* __restrict__ acc2, int renorm,
                                                                     float floor, float temp, float* __restrict__ p,
                                                                     double* p_stats, void* ws) {
  const float sum = (float)acc2[0];
  const float maxc = fmaxf((float)acc2[1] / sum, floor);
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float v = p2[i];
    if (renorm) v = fmaxf(v / sum, floor) / maxc;
    if (temp != 1.0f) v = powf(v, temp);
    p[i] = v;
    s += (double)((v != v) ? 1e-6f : v);  // cost_norm sanitises NaN before summing
  }
  const int kinds[1] = {RED_SUM};
  double vals[1] = {s};
  grid_reduce<1>(kinds, vals, ws_misc_partials(ws), ws_misc_counter(ws, 1), blockIdx.x, gridDim.x, p_stats);
}

// dynamics, barrier, rollout and adjoint device code lives in klerg_dyn.cuh

// One CTA per candidate (rollout_block): controls, trajectory and linearisation staged in shared memory.
constexpr int RO_THREADS = 64;

__host__ __device__ inline size_t rollout_smem_floats(int kind, int S, int A, int64_t H, bool want_dbarr, bool want_P) {
  size_t n = (size_t)H * A + (size_t)(H + 1) * S + 34;
  if (want_dbarr) n += (size_t)H * S;
  if (kind == KLERG_DYN_ROLL) {
    n += rollout_rot_floats(1, (int)H);
    if (want_P) n += (size_t)H * A * A;
  }
  return n;
}

__global__ void __launch_bounds__(RO_THREADS) rollout_kernel(DynDev d, BarDev bar, const float* __restrict__ x0,
                                                             const float* __restrict__ R0,
                                                             const float* __restrict__ u, int64_t B, int64_t H64,
                                                             float* __restrict__ traj,
                                                             float* __restrict__ barrier_sum,
                                                             float* __restrict__ dbarr, float* __restrict__ P,
                                                             float* __restrict__ R_out) {
  extern __shared__ float sh[];
  const int H = (int)H64, S = d.S, a = d.A;
  const bool roll = d.kind == KLERG_DYN_ROLL;
  const int64_t b = blockIdx.x;
  float* s_u = sh;
  float* s_traj = s_u + H * a;
  float* s_red = s_traj + (H + 1) * S;   // 32 + 1 (bsum) + pad
  float* s_dbarr = s_red + 34;
  float* s_rot = s_dbarr + (dbarr ? H * S : 0);
  float* s_P = s_rot + (roll ? rollout_rot_floats(1, H) : 0);
  for (int e = threadIdx.x; e < H * a; e += blockDim.x) s_u[e] = u[b * H * a + e];
  __syncthreads();
  rollout_block(d, bar, x0, R0, s_u, 1, H, s_traj, dbarr ? s_dbarr : nullptr, (P && roll) ? s_P : nullptr, s_rot, s_red,
                s_red + 32, R_out ? R_out + b * 9 : nullptr);
  for (int e = threadIdx.x; e < (H + 1) * S; e += blockDim.x) traj[b * (H + 1) * S + e] = s_traj[e];
  if (dbarr)
    for (int e = threadIdx.x; e < H * S; e += blockDim.x) dbarr[b * H * S + e] = s_dbarr[e];
  if (P)
    for (int e = threadIdx.x; e < H * a * a; e += blockDim.x)
      P[b * H * a * a + e] = roll ? s_P[e] : (((e % (a * a)) / a == e % a) ? 0.8f : 0.f);
  if (threadIdx.x == 0 && barrier_sum) barrier_sum[b] = s_red[32];
}

__global__ void barrier_eval_kernel(BarDev bar, const float* __restrict__ x, int64_t T, int S, float* __restrict__ value,
                                    float* __restrict__ grad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float row[KLERG_MAX_S], g[KLERG_MAX_S];
  for (int i = 0; i < S; ++i) row[i] = x[t * S + i];
  if (value) value[t] = barrier_value(bar, row);
  if (grad) {
    barrier_grad(bar, row, S, g);
    for (int i = 0; i < S; ++i) grad[t * S + i] = g[i];
  }
}

struct TiltDev {
  int on, r_idx, p_idx, w_idx, has_map;
  float w_lo, w_hi, lim, power, weight;
  float rot_lo[2], rot_hi[2], ang_lo[2], ang_hi[2];
};

// VelocityBarrier (limits relative to x_ref) and TiltBarrierFunction (tilt-dependent yaw limits + tilt term)
__global__ void barrier_ext_kernel(BarDev bar, TiltDev tl, const float* __restrict__ x, const float* __restrict__ x_ref,
                                   int64_t T, int S, float* __restrict__ value, float* __restrict__ grad,
                                   float* __restrict__ tilt_out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float row[KLERG_MAX_S], g[KLERG_MAX_S];
  for (int i = 0; i < S; ++i) row[i] = x[t * S + i];
  if (x_ref)
    for (int i = 0; i < bar.n; ++i) {
      bar.lo[i] += x_ref[t * S + i];
      bar.hi[i] += x_ref[t * S + i];
    }
  float tv = 0.f, gr = 0.f, gp = 0.f;
  if (tl.on) {
    float r = row[tl.r_idx], p = row[tl.p_idx];
    if (tl.has_map) {
      r = affine_map(r, tl.rot_lo[0], tl.rot_hi[0], tl.ang_lo[0], tl.ang_hi[0]);
      p = affine_map(p, tl.rot_lo[1], tl.rot_hi[1], tl.ang_lo[1], tl.ang_hi[1]);
    }
    float sr, cr, sp, cp;
    sincosf(r, &sr, &cr);
    sincosf(p, &sp, &cp);
    const float tilt = acosf(cr * cp);
    bar.lo[tl.w_idx] = tilt / 3.14159265358979323846f * tl.w_lo;
    bar.hi[tl.w_idx] = tilt / 3.14159265358979323846f * tl.w_hi;
    if (tilt <= tl.lim) {
      const float d = tilt - tl.lim;
      tv = tl.weight * powi_or_f(d, tl.power);
      const float dd = tl.power * tl.weight * powi_or_f(d, tl.power - 1.f) / sqrtf(-cp * cp * cr * cr + 1.f);
      gr = dd * sr * cp;
      gp = dd * sp * cr;
    }
    if (tilt_out) tilt_out[t] = tilt;
  }
  if (value) value[t] = tv + barrier_value(bar, row);
  if (grad) {
    barrier_grad(bar, row, S, g);
    if (tl.on) {
      g[tl.r_idx] += gr;
      g[tl.p_idx] += gp;
    }
    for (int i = 0; i < S; ++i) grad[t * S + i] = g[i];
  }
}

// ---------------------------------------------------------------------------
// adjoint sweep (klerg.py:433-450, 590-593): stage inputs in shared memory, then adjoint_block
// ---------------------------------------------------------------------------
struct AdjArgs {
  DynDev d;
  KernelDev k;
  int64_t H;
  const double* grad_part;
  int world;
  const float* dbarr;
  const float* P;
  const float* traj;
  const float* u;
  AdjParams ap;
  float* dgdx;
  float* du;
  float* djdlam;
  float* u_star;
  int64_t gp_stride;  // doubles between the gradient blocks of consecutive targets (one CTA per target)
};

__global__ void __launch_bounds__(256) adjoint_kernel(const AdjArgs a0) {
  // one CTA per belief target: same trajectory / linearisation / controls, its own gradient and outputs
  AdjArgs a = a0;
  {
    const int64_t kt = blockIdx.x;
    a.grad_part += kt * a.gp_stride;
    a.dgdx += kt * a.H * a.d.S;
    a.du += kt * a.H * a.d.A;
    a.djdlam += kt * a.H;
    a.u_star += kt * a.H * a.d.A;
  }
  extern __shared__ float sh[];  // g[H][S] | P[H][A*A] (optional) | traj[H][S] (SPEED) | u[H][A] | scratch
  const int S = a.d.S, A = a.d.A, D = a.k.D;
  const int64_t H = a.H;
  const bool speed = a.d.kind == KLERG_DYN_SPEED;
  float* sg = sh;
  float* sP = sg + H * S;
  float* straj = sP + (a.P ? H * A * A : 0);
  float* su = straj + (speed ? H * S : 0);
  float* s_scr = su + H * A;
  // phase 1 (all threads): g = dgdx - dbarr staged in smem, dgdx[H][S] written out
  for (int64_t e = threadIdx.x; e < H * S; e += blockDim.x) sg[e] = 0.f;
  __syncthreads();
  for (int64_t e = threadIdx.x; e < H * D; e += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < a.world; ++r) s += a.grad_part[(size_t)r * H * D + e];
    sg[(e / D) * S + a.k.explr[e % D]] = (float)s;
  }
  __syncthreads();
  for (int64_t e = threadIdx.x; e < H * S; e += blockDim.x) {
    const float g = sg[e];
    a.dgdx[e] = g;
    sg[e] = g - a.dbarr[e];
  }
  if (a.P)
    for (int64_t e = threadIdx.x; e < H * A * A; e += blockDim.x) sP[e] = a.P[e];
  for (int64_t e = threadIdx.x; e < H * A; e += blockDim.x) su[e] = a.u[e];
  if (speed)
    for (int64_t e = threadIdx.x; e < H * S; e += blockDim.x) straj[e] = a.traj[e];
  __syncthreads();
  adjoint_block(a.d, a.ap, (int)H, sg, a.P ? sP : nullptr, straj, su, s_scr, a.du, a.djdlam, a.u_star);
}

struct CombineArgs {
  int n, world;
  int kind[8];
};

__global__ void combine_blocks_kernel(CombineArgs c, const double* __restrict__ blocks, double* __restrict__ out) {
  const int q = threadIdx.x;
  if (q >= c.n) return;
  double v = red_identity(c.kind[q]);
  for (int r = 0; r < c.world; ++r) v = red_combine(c.kind[q], v, blocks[(size_t)r * c.n + q]);
  out[q] = v;
}

// expo = {mean_i spread'_i, min p, max spread} from acc = {smax, sum_inside, n_outside, min p}
__global__ void target_exponent_kernel(const double* __restrict__ acc, int64_t N_total, double* __restrict__ expo) {
  if (threadIdx.x == 0) {
    const float smax = (float)acc[0];
    const float mean = (float)(((double)((float)acc[1] / smax) + acc[2]) / (double)N_total);
    expo[0] = (double)mean;
    expo[1] = acc[3];
    expo[2] = acc[0];
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ table, int S, const int64_t* __restrict__ idx, int64_t M,
                                   float* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * S; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = e / S;
    out[e] = table[idx[m] * S + (e - m * S)];
  }
}

bool make_dyn(const klerg_dyn_spec* s, DynDev& d) {
  if (!s) { set_error("dyn spec is null"); return false; }
  if (s->kind < 0 || s->kind > 3 || s->S < 1 || s->S > KLERG_MAX_S || s->A < 1 || s->A > KLERG_MAX_A) {
    set_error("dyn spec out of range (kind=%d S=%d A=%d)", s->kind, s->S, s->A);
    return false;
  }
  const int need = s->kind == KLERG_DYN_SINGLE ? s->A : (s->kind == KLERG_DYN_SPEED ? 3 * s->A : 2 * s->A);
  if (s->S != need) { set_error("dyn spec: S=%d inconsistent with kind %d, A=%d", s->S, s->kind, s->A); return false; }
  d.kind = s->kind; d.S = s->S; d.A = s->A; d.dt = s->dt; d.has_map = s->has_ang_map;
  for (int k = 0; k < 3; ++k) {
    d.rpw[k] = s->rpw[k];
    d.rot_lo[k] = s->rot_lo[k]; d.rot_hi[k] = s->rot_hi[k];
    d.ang_lo[k] = s->ang_lo[k]; d.ang_hi[k] = s->ang_hi[k];
    if (s->kind == KLERG_DYN_ROLL && (s->rpw[k] < 0 || s->rpw[k] >= s->A)) { set_error("dyn spec: rpw index out of range"); return false; }
  }
  return true;
}

bool make_bar(const klerg_barrier_spec* s, BarDev& b) {
  memset(&b, 0, sizeof(b));
  if (!s) return true;  // NoBarrier
  if (s->n < 0 || s->n > KLERG_MAX_S) { set_error("barrier spec: n out of range"); return false; }
  b.n = s->n;
  for (int i = 0; i < s->n; ++i) { b.lo[i] = s->lo[i]; b.hi[i] = s->hi[i]; b.w[i] = s->weight[i]; b.pw[i] = s->power[i]; }
  return true;
}

static bool make_lim(int D, const float* lo, const float* hi, LimDev& L) {
  if (D < 1 || D > KLERG_MAX_D) { set_error("target: D out of range"); return false; }
  L.D = D;
  for (int d = 0; d < D; ++d) { L.lo[d] = lo ? lo[d] : -INFINITY; L.hi[d] = hi ? hi[d] : INFINITY; }
  return true;
}

}  // namespace klerg

using namespace klerg;

extern "C" const char* klerg_last_error(void) { return g_err; }
extern "C" int klerg_abi_version(void) { return 6; }
extern "C" long long klerg_launch_count(void) { return g_launches; }

// ---- peak-rate microbenchmarks (roofline denominators, used by bench.py only) ----
namespace klerg {
template <int KIND>
__global__ void __launch_bounds__(256) peak_kernel(int iters, float seed, float* out) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + 1e-3f * (float)(threadIdx.x + i);
  const float m = 0.999f, c = 1e-4f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) a[i] = fmaf(a[i], m, c);       // FFMA
        else a[i] = ex2_approx(a[i] * -0.5f);          // MUFU.EX2 (+1 FMUL)
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;  // keep the chain alive
}
// FFMA2 (fma.rn.f32x2): two fp32 lanes per instruction, three register operands
__global__ void __launch_bounds__(256) peak_kernel_packed(int iters, float seed, float* out) {
  unsigned long long a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x = seed + 1e-3f * (float)(threadIdx.x + i);
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b[i]) : "f"(0.999f + 1e-6f * (float)i));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c[i]) : "f"(1e-4f + 1e-6f * (float)i));
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b[i]), "l"(c[i]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    s += lo + hi;
  }
  if (s == 123.456f) out[0] = s;
}
}  // namespace klerg

extern "C" int klerg_peak_probe(int kind, int iters, int blocks, float* out, void* stream) {
  if (kind == 0) klerg::peak_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.5f, out);
  else if (kind == 1) klerg::peak_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.5f, out);
  else if (kind == 2) klerg::peak_kernel_packed<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.5f, out);
  else { set_error("peak_probe: kind 0 (FFMA), 1 (EX2) or 2 (FFMA2)"); return -1; }
  return check_launch("peak_kernel");
}

extern "C" int klerg_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("device_info: %s", cudaGetErrorString(e)); return -4; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { set_error("device_info: %s", cudaGetErrorString(e)); return -4; }
  if (sms) *sms = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  return 0;
}

extern "C" size_t klerg_workspace_bytes(int64_t G) {
  if (G < 1) G = 1;
  return HEAD_BYTES + (size_t)G * SEG_BYTES;
}

extern "C" int klerg_vector_stats(const float* x, int64_t N, double* stats, void* ws, void* stream) {
  vector_stats_kernel<<<ew_blocks(N), EW_THREADS, 0, (cudaStream_t)stream>>>(x, N, stats, ws);
  return check_launch("vector_stats_kernel");
}

extern "C" int klerg_renormalize(const float* x, int64_t N, float floor, float* out, void* ws, void* stream) {
  double* stats = ws_misc_partials(ws) + (size_t)MAXBLK * 4;  // tail of the misc area
  vector_stats_kernel<<<ew_blocks(N), EW_THREADS, 0, (cudaStream_t)stream>>>(x, N, stats, ws);
  renorm_apply_kernel<<<ew_blocks(N), EW_THREADS, 0, (cudaStream_t)stream>>>(x, N, stats, floor, out);
  return check_launch("renormalize");
}

extern "C" int klerg_renormalize_with_stats(const float* x, int64_t N, const double* stats, float floor, float* out,
                                            void* stream) {
  renorm_apply_kernel<<<ew_blocks(N), EW_THREADS, 0, (cudaStream_t)stream>>>(x, N, stats, floor, out);
  return check_launch("renorm_apply_kernel");
}

extern "C" int klerg_cost_norm(float* x, int64_t N, void* ws, void* stream) {
  double* stats = ws_misc_partials(ws) + (size_t)MAXBLK * 4;
  cost_norm_pass1<<<ew_blocks(N), EW_THREADS, 0, (cudaStream_t)stream>>>(x, N, stats, ws);
  cost_norm_pass2<<<ew_blocks(N), EW_THREADS, 0, (cudaStream_t)stream>>>(x, N, stats);
  return check_launch("cost_norm");
}

extern "C" int klerg_kl_cost_partial(const float* v, int64_t v_stride, int64_t G, int64_t N, const double* totals,
                                     int world, const float* p, float floor, double* kl_part, void* ws,
                                     void* stream) {
  if (G < 1) { set_error("kl_cost_partial: G < 1"); return -1; }
  int gx = ew_blocks(N);
  if (G > 1) {
    // many candidates: keep the total CTA count near 8 waves
    int64_t cap = ((int64_t)sm_count() * 8 + G - 1) / G;
    if (cap < 1) cap = 1;
    if (gx > cap) gx = (int)cap;
  }
  const unsigned gy = (unsigned)(G > 65535 ? 65535 : G);
  kl_cost_partial_kernel<<<dim3(gx, gy), EW_THREADS, 0, (cudaStream_t)stream>>>(v, v_stride, G, N, totals, world, p,
                                                                                 floor, kl_part, ws);
  return check_launch("kl_cost_partial_kernel");
}

extern "C" int klerg_kl_cost_final(const double* kl_part, int world, int64_t G, const double* p_stats,
                                   const float* barrier_sum, float* cost, void* stream) {
  kl_cost_final_kernel<<<(unsigned)((G + 127) / 128), 128, 0, (cudaStream_t)stream>>>(kl_part, world, G, p_stats,
                                                                                      barrier_sum, cost);
  return check_launch("kl_cost_final_kernel");
}

extern "C" int klerg_target_stage1(const float* samples, int32_t D, int64_t N, const float* lim_lo,
                                   const float* lim_hi, const float* spread, const float* p, double* acc, void* ws,
                                   void* stream) {
  LimDev L;
  if (!make_lim(D, lim_lo, lim_hi, L)) return -1;
  target_stage1_kernel<<<ew_blocks(N), EW_THREADS, 0, (cudaStream_t)stream>>>(samples, L, N, spread, p, acc, ws);
  return check_launch("target_stage1_kernel");
}

extern "C" int klerg_combine_blocks(const double* blocks, int world, int n, const int* kinds, double* out,
                                    void* stream) {
  if (n < 1 || n > 8 || world < 1) { set_error("combine_blocks: n in 1..8, world >= 1"); return -1; }
  CombineArgs c{};
  c.n = n; c.world = world;
  for (int q = 0; q < n; ++q) {
    if (kinds[q] < 0 || kinds[q] > 2) { set_error("combine_blocks: bad kind"); return -1; }
    c.kind[q] = kinds[q];
  }
  combine_blocks_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c, blocks, out);
  return check_launch("combine_blocks_kernel");
}

extern "C" int klerg_target_exponent(const double* acc, int64_t N_total, double* expo, void* stream) {
  if (N_total < 1) { set_error("target_exponent: N_total < 1"); return -1; }
  target_exponent_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, N_total, expo);
  return check_launch("target_exponent_kernel");
}

extern "C" int klerg_target_stage2(int mode, const float* samples, int32_t D, int64_t N, const float* lim_lo,
                                   const float* lim_hi, const float* spread, const float* p, const double* expo,
                                   float* p2, double* acc, void* ws, void* stream) {
  LimDev L;
  if (!make_lim(D, lim_lo, lim_hi, L)) return -1;
  if (mode < 0 || mode > 2) { set_error("target_stage2: bad mode"); return -1; }
  target_stage2_kernel<<<ew_blocks(N), EW_THREADS, 0, (cudaStream_t)stream>>>(mode, samples, L, N, spread, p, expo,
                                                                               p2, acc, ws);
  return check_launch("target_stage2_kernel");
}

extern "C" int klerg_target_stage3(const float* p2, int64_t N, const double* acc2, int renorm, float floor,
                                   float temp, float* p, double* p_stats, void* ws, void* stream) {
  target_stage3_kernel<<<ew_blocks(N), EW_THREADS, 0, (cudaStream_t)stream>>>(p2, N, acc2, renorm, floor, temp, p,
                                                                               p_stats, ws);
  return check_launch("target_stage3_kernel");
}

extern "C" int klerg_rollout(const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar, const float* x0,
                             const float* R0, const float* u, int64_t B, int64_t H, float* traj,
                             float* barrier_sum, float* dbarr, float* P, float* R_out, void* stream) {
  DynDev d;
  BarDev b;
  if (!make_dyn(dyn, d) || !make_bar(bar, b)) return -1;
  if (B < 1 || H < 0) { set_error("rollout: bad sizes"); return -1; }
  if (H > KLERG_MAX_H) { set_error("rollout: H > KLERG_MAX_H"); return -1; }
  const size_t smem = sizeof(float) * rollout_smem_floats(d.kind, d.S, d.A, H, dbarr != nullptr, P != nullptr);
  if (smem > 48 * 1024) {
    static size_t raised = 0;
    if (smem > 200 * 1024) { set_error("rollout: horizon too long for shared-memory staging"); return -1; }
    if (smem > raised) {
      cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      raised = smem;
    }
  }
  rollout_kernel<<<(unsigned)B, RO_THREADS, smem, (cudaStream_t)stream>>>(
      d, b, x0, R0, u, B, H, traj, barrier_sum, dbarr, P, R_out);
  return check_launch("rollout_kernel");
}

extern "C" int klerg_barrier_eval(const klerg_barrier_spec* bar, const float* x, int64_t T, int32_t S, float* value,
                                  float* grad, void* stream) {
  BarDev b;
  if (!make_bar(bar, b)) return -1;
  if (S < 1 || S > KLERG_MAX_S) { set_error("barrier_eval: S out of range"); return -1; }
  if (T < 1) return 0;
  barrier_eval_kernel<<<(unsigned)((T + 127) / 128), 128, 0, (cudaStream_t)stream>>>(b, x, T, S, value, grad);
  return check_launch("barrier_eval_kernel");
}

extern "C" int klerg_barrier_eval_ext(const klerg_barrier_spec* bar, const float* x, const float* x_ref, const klerg_tilt_spec* tilt,
                                      int64_t T, int32_t S, float* value, float* grad, float* tilt_out, void* stream) {
  BarDev b;
  if (!make_bar(bar, b)) return -1;
  if (S < 1 || S > KLERG_MAX_S) { set_error("barrier_eval_ext: S out of range"); return -1; }
  TiltDev tl{};
  if (tilt) {
    if (tilt->r_idx < 0 || tilt->r_idx >= S || tilt->p_idx < 0 || tilt->p_idx >= S || tilt->w_idx < 0 || tilt->w_idx >= b.n) {
      set_error("barrier_eval_ext: tilt indices out of range");
      return -1;
    }
    tl.on = 1; tl.r_idx = tilt->r_idx; tl.p_idx = tilt->p_idx; tl.w_idx = tilt->w_idx; tl.has_map = tilt->has_map;
    tl.w_lo = tilt->w_lo; tl.w_hi = tilt->w_hi; tl.lim = tilt->tilt_lim; tl.power = tilt->power; tl.weight = tilt->weight;
    for (int k = 0; k < 2; ++k) {
      tl.rot_lo[k] = tilt->rot_lo[k]; tl.rot_hi[k] = tilt->rot_hi[k];
      tl.ang_lo[k] = tilt->ang_lo[k]; tl.ang_hi[k] = tilt->ang_hi[k];
    }
  }
  if (T < 1) return 0;
  if (!x) { set_error("barrier_eval_ext: null input"); return -1; }
  barrier_ext_kernel<<<(unsigned)((T + 127) / 128), 128, 0, (cudaStream_t)stream>>>(b, tl, x, x_ref, T, S, value, grad, tilt_out);
  return check_launch("barrier_ext_kernel");
}

// ---------------------------------------------------------------------------
// State-feedback default policies (default_policies.py:53-119): the closed loop is serial in t (u_t depends on x_t),
// so one warp steps it; H <= 256 steps of a <= 24-state model are a few microseconds.
// ---------------------------------------------------------------------------
struct PolicyDev {
  int kind, use_u;
  float weight;
  float K[KLERG_MAX_A * KLERG_MAX_S];
};

__global__ void __launch_bounds__(32) policy_rollout_kernel(const DynDev d, const PolicyDev pol, const float* __restrict__ x0,
                                                            const float* __restrict__ R0, const float* __restrict__ u_in,
                                                            int H, float* __restrict__ u_eff, float* __restrict__ dmudx) {
  __shared__ float x[KLERG_MAX_S], u[KLERG_MAX_A], Rm[9];
  const int lane = threadIdx.x, S = d.S, A = d.A;
  const bool single = d.kind == KLERG_DYN_SINGLE, roll = d.kind == KLERG_DYN_ROLL;
  const float dt = d.dt;
  if (lane < S) x[lane] = x0[lane];
  __syncwarp();
  if (roll && lane == 0) {
    if (R0) {
      for (int i = 0; i < 9; ++i) Rm[i] = R0[i];
    } else {
      float rot[3];
      for (int k = 0; k < 3; ++k) {
        rot[k] = x[d.rpw[k]];
        if (d.has_map) rot[k] = affine_map(rot[k], d.rot_lo[k], d.rot_hi[k], d.ang_lo[k], d.ang_hi[k]);
      }
      euler_xyz_to_matrix(rot, Rm);
    }
  }
  __syncwarp();
  for (int t = 0; t < H; ++t) {
    // u_t = policy(x_t), dmudx_t
    for (int e = lane; e < A * S; e += 32) dmudx[(size_t)t * A * S + e] = pol.kind == KLERG_POLICY_LQR ? -pol.K[e] : 0.f;
    __syncwarp();
    if (lane < A) {
      float ui;
      if (pol.kind == KLERG_POLICY_LQR) {
        ui = 0.f;
        for (int j = 0; j < S; ++j) ui = fmaf(pol.K[lane * S + j], x[j], ui);  // (K @ x)_i, then negated
        ui = -ui;
      } else {
        ui = pol.use_u ? u_in[t * A + lane] : 0.f;
        if (!single) {
          const float p = x[lane], v = x[A + lane];
          if ((p >= 1.f && v > 0.f) || (p <= -1.f && v < 0.f)) {
            ui = -pol.weight * v;
            dmudx[(size_t)t * A * S + lane * S + A + lane] = -pol.weight;
          }
        }
      }
      u[lane] = ui;
      u_eff[t * A + lane] = ui;
    }
    __syncwarp();
    // one exact RK4 step of the integrator model (dynamics.py:7-13,58-65; A nilpotent) + the rotation update
    float E[9], Rn[9];
    if (roll && lane == 0) {
      float w3[3];
      for (int k = 0; k < 3; ++k) w3[k] = x[A + d.rpw[k]];  // pre-step angular velocity (dynamics.py:270-272)
      rodrigues(w3, dt, E);
      matmul3(E, Rm, Rn);
    }
    __syncwarp();
    if (lane < A) {
      // The policy branches on the state (x_i >= 1, v_i > 0), and BarrierPush's own control u = -5 v with dt = 0.2
      // lands the velocity EXACTLY on 0 in the reference: the step is evaluated in the reference's operation order
      // (rk4_integrate, dynamics.py:7-13: k = dt f, x + (1/6)(k1 + 2 k2 + 2 k3 + k4), no contraction), not in the
      // closed form of the open-loop rollout, so that such ties fall the same way.
      const float sixth = 0.16666667163372040f;  // float32(1/6.)
      const float ku = __fmul_rn(dt, u[lane]);   // velocity rate = u at every stage
      const float su = __fadd_rn(__fadd_rn(__fadd_rn(ku, __fmul_rn(2.f, ku)), __fmul_rn(2.f, ku)), ku);
      if (single) {
        x[lane] = __fadd_rn(x[lane], __fmul_rn(sixth, su));
      } else {
        const float v = x[A + lane];
        const float k1 = __fmul_rn(dt, __fmul_rn(0.8f, v));
        const float k2 = __fmul_rn(dt, __fmul_rn(0.8f, __fadd_rn(v, __fmul_rn(ku, 0.5f))));
        const float k4 = __fmul_rn(dt, __fmul_rn(0.8f, __fadd_rn(v, ku)));
        const float sp = __fadd_rn(__fadd_rn(__fadd_rn(k1, __fmul_rn(2.f, k2)), __fmul_rn(2.f, k2)), k4);
        x[lane] = __fadd_rn(x[lane], __fmul_rn(sixth, sp));
        x[A + lane] = __fadd_rn(v, __fmul_rn(sixth, su));
      }
    }
    __syncwarp();
    if (roll && lane == 0) {
      float rot[3];
      wrapped_euler_xyz(Rn, rot);
      for (int k = 0; k < 3; ++k) {
        float v = rot[k];
        if (d.has_map) v = affine_map(v, d.ang_lo[k], d.ang_hi[k], d.rot_lo[k], d.rot_hi[k]);
        x[d.rpw[k]] = v;
      }
      for (int i = 0; i < 9; ++i) Rm[i] = Rn[i];
    }
    __syncwarp();
  }
}

struct AdjPolicyArgs {
  DynDev d;
  AdjParams ap;
  int H;
  const float *dgdx, *dbarr, *P, *dmudx, *u;
  float *du, *djdlam, *u_star;
};

// rho_H = 0; t = H-1..0: rho <- rk4(rho' = g_t - (A_t + B dmudx_t)^T rho, step -dt) (klerg.py:433-450, 590-593 with
// dynamics.py:7-13), lane j holds rho_j; du_t = -Rinv B^T rho, djdlam_t = rho B du_t, u* as in adjoint_block.
__global__ void __launch_bounds__(32) adjoint_policy_kernel(const AdjPolicyArgs a) {
  __shared__ float M[KLERG_MAX_S * KLERG_MAX_S];  // closed-loop A_t + B dmudx_t, row-major [S][S]
  __shared__ float s_du[KLERG_MAX_A], s_btr[KLERG_MAX_A];
  const int lane = threadIdx.x, S = a.d.S, A = a.d.A;
  const bool single = a.d.kind == KLERG_DYN_SINGLE;
  const float h = -a.d.dt;
  float rho = 0.f;
  for (int t = a.H - 1; t >= 0; --t) {
    for (int e = lane; e < S * S; e += 32) {
      const int r = e / S, c = e - r * S;
      float m = 0.f;
      if (!single) {
        if (r < A && c >= A) m = a.P ? a.P[(size_t)t * A * A + r * A + (c - A)] : (c - A == r ? 0.8f : 0.f);
        if (r >= A) m = a.dmudx[(size_t)t * A * S + (r - A) * S + c];  // B = [0; I]
      } else {
        m = a.dmudx[(size_t)t * A * S + e];  // A = 0, B = I
      }
      M[e] = m;
    }
    __syncwarp();
    const float g = lane < S ? a.dgdx[t * S + lane] - a.dbarr[t * S + lane] : 0.f;
    auto f = [&](float r) {  // g - M^T r, every lane its own component
      float acc = 0.f;
      for (int i = 0; i < S; ++i) acc = fmaf(M[i * S + (lane < S ? lane : 0)], __shfl_sync(0xffffffffu, r, i), acc);
      return g - acc;
    };
    const float k1 = h * f(rho);
    const float k2 = h * f(rho + k1 / 2.f);
    const float k3 = h * f(rho + k2 / 2.f);
    const float k4 = h * f(rho + k3);
    rho = lane < S ? rho + (1.f / 6.f) * (k1 + 2.f * k2 + 2.f * k3 + k4) : 0.f;
    // B^T rho: the velocity half (DOUBLE / ROLL) or rho itself (SINGLE)
    const float btr = __shfl_sync(0xffffffffu, rho, single ? (lane < A ? lane : 0) : (lane < A ? A + lane : 0));
    if (lane < A) {
      const float dui = -a.ap.rinv[lane] * btr;
      s_du[lane] = dui;
      s_btr[lane] = btr;
      a.du[t * A + lane] = dui;
      const float us = a.u[t * A + lane] + a.ap.alpha * dui;
      a.u_star[t * A + lane] = a.ap.sat > 0.f ? tanhf(us / a.ap.sat) * a.ap.chi[lane] : fminf(fmaxf(us, a.ap.clo[lane]), a.ap.chi[lane]);
    }
    __syncwarp();
    if (lane == 0) {
      float dj = 0.f;
      for (int i = 0; i < A; ++i) dj += s_btr[i] * s_du[i];
      a.djdlam[t] = dj;
    }
    __syncwarp();
  }
}

static bool make_policy(const klerg_policy_spec* s, const DynDev& d, PolicyDev& p) {
  if (!s) { set_error("policy spec is null"); return false; }
  if (s->kind != KLERG_POLICY_LQR && s->kind != KLERG_POLICY_BARRIER_PUSH) { set_error("policy spec: unknown kind %d", s->kind); return false; }
  if (d.kind == KLERG_DYN_SPEED) { set_error("state-feedback policies: the speed-state model is not supported"); return false; }
  p.kind = s->kind; p.use_u = s->use_u; p.weight = s->weight;
  for (int e = 0; e < KLERG_MAX_A * KLERG_MAX_S; ++e) p.K[e] = e < d.A * d.S ? s->K[e] : 0.f;
  return true;
}

extern "C" int klerg_policy_rollout(const klerg_dyn_spec* dyn, const klerg_policy_spec* pol, const float* x0, const float* R0,
                                    const float* u_in, int64_t H, float* u_eff, float* dmudx, void* stream) {
  DynDev d;
  PolicyDev p;
  if (!make_dyn(dyn, d) || !make_policy(pol, d, p)) return -1;
  if (H < 1 || H > KLERG_MAX_H) { set_error("policy_rollout: H out of range"); return -1; }
  if (!x0 || !u_eff || !dmudx || (p.kind == KLERG_POLICY_BARRIER_PUSH && p.use_u && !u_in)) { set_error("policy_rollout: null argument"); return -1; }
  policy_rollout_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d, p, x0, R0, u_in, (int)H, u_eff, dmudx);
  return check_launch("policy_rollout_kernel");
}

extern "C" int klerg_adjoint_policy(const klerg_dyn_spec* dyn, int64_t H, const float* dgdx, const float* dbarr, const float* P,
                                    const float* dmudx, const float* u, const float* Rinv_diag, float alpha,
                                    const float* ctrl_lo, const float* ctrl_hi, float* du, float* djdlam, float* u_star,
                                    void* stream) {
  AdjPolicyArgs a{};
  if (!make_dyn(dyn, a.d)) return -1;
  if (a.d.kind == KLERG_DYN_SPEED) { set_error("adjoint_policy: the speed-state model is not supported"); return -1; }
  if (H < 1 || H > KLERG_MAX_H) { set_error("adjoint_policy: H out of range"); return -1; }
  if (!dgdx || !dbarr || !dmudx || !u || !du || !djdlam || !u_star || !Rinv_diag || !ctrl_lo || !ctrl_hi) { set_error("adjoint_policy: null argument"); return -1; }
  a.H = (int)H; a.dgdx = dgdx; a.dbarr = dbarr; a.P = P; a.dmudx = dmudx; a.u = u; a.du = du; a.djdlam = djdlam; a.u_star = u_star;
  a.ap.alpha = alpha;
  a.ap.sat = 1e-3f * (float)g_saturate_milli;
  for (int i = 0; i < a.d.A; ++i) { a.ap.rinv[i] = Rinv_diag[i]; a.ap.clo[i] = ctrl_lo[i]; a.ap.chi[i] = ctrl_hi[i]; }
  adjoint_policy_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
  return check_launch("adjoint_policy_kernel");
}

static int launch_adjoint(const klerg_dyn_spec* dyn, const klerg_kernel_spec* k, int64_t H, const double* grad_part,
                          int world, int64_t K, int64_t gp_stride, const float* dbarr, const float* P, const float* traj,
                          const float* u, const float* Rinv_diag, float alpha, const float* ctrl_lo, const float* ctrl_hi,
                          float* dgdx, float* du, float* djdlam, float* u_star, void* stream) {
  AdjArgs a{};
  if (!make_dyn(dyn, a.d) || !make_kernel_dev(k, a.k)) return -1;
  if (H < 1 || H > KLERG_MAX_H) { set_error("adjoint: H out of range"); return -1; }
  if (K < 1 || K > 65535) { set_error("adjoint: K out of range"); return -1; }
  a.H = H; a.grad_part = grad_part; a.world = world; a.dbarr = dbarr; a.P = P; a.traj = traj; a.u = u;
  a.ap.alpha = alpha; a.dgdx = dgdx; a.du = du; a.djdlam = djdlam; a.u_star = u_star; a.gp_stride = gp_stride;
  for (int i = 0; i < a.d.A; ++i) { a.ap.rinv[i] = Rinv_diag[i]; a.ap.clo[i] = ctrl_lo[i]; a.ap.chi[i] = ctrl_hi[i]; }
  a.ap.sat = 1e-3f * (float)g_saturate_milli;
  const bool speed = a.d.kind == KLERG_DYN_SPEED;
  const size_t smem = sizeof(float) * ((size_t)H * (a.d.S + (P ? a.d.A * a.d.A : 0) + (speed ? a.d.S : 0) + a.d.A) +
                                       adjoint_scratch_floats((int)H, a.d.A));
  if (smem > 48 * 1024) {
    static bool raised = false;
    if (!raised) {
      cudaFuncSetAttribute(adjoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      raised = true;
    }
    if (smem > 200 * 1024) { set_error("adjoint: horizon too long for shared-memory staging"); return -1; }
  }
  adjoint_kernel<<<(unsigned)K, 256, smem, (cudaStream_t)stream>>>(a);
  return check_launch("adjoint_kernel");
}

extern "C" int klerg_adjoint(const klerg_dyn_spec* dyn, const klerg_kernel_spec* k, int64_t H,
                             const double* grad_part, int world, const float* dbarr, const float* P,
                             const float* traj, const float* u, const float* Rinv_diag, float alpha,
                             const float* ctrl_lo, const float* ctrl_hi, float* dgdx, float* du, float* djdlam,
                             float* u_star, void* stream) {
  return launch_adjoint(dyn, k, H, grad_part, world, 1, 0, dbarr, P, traj, u, Rinv_diag, alpha, ctrl_lo, ctrl_hi, dgdx, du,
                        djdlam, u_star, stream);
}

extern "C" int klerg_adjoint_targets(const klerg_dyn_spec* dyn, const klerg_kernel_spec* k, int64_t H, int64_t K,
                                     const double* grad_parts, int world, const float* dbarr, const float* P,
                                     const float* traj, const float* u, const float* Rinv_diag, float alpha,
                                     const float* ctrl_lo, const float* ctrl_hi, float* dgdx, float* du,
                                     float* djdlam, float* u_star, void* stream) {
  if (!k) { set_error("kernel spec is null"); return -1; }
  return launch_adjoint(dyn, k, H, grad_parts, world, K, (int64_t)world * H * k->D, dbarr, P, traj, u, Rinv_diag, alpha,
                        ctrl_lo, ctrl_hi, dgdx, du, djdlam, u_star, stream);
}

extern "C" int klerg_gather_rows(const float* table, int32_t S, const int64_t* idx, int64_t M, float* out,
                                 void* stream) {
  if (M < 1) return 0;
  int64_t blocks = (M * S + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(table, S, idx, M, out);
  return check_launch("gather_rows_kernel");
}

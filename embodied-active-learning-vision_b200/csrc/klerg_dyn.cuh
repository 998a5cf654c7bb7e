// Planner-model device code shared by the stand-alone kernels (klerg_planner.cu) and the
// fused eval kernels (klerg_fused.cu): integrator models (dynamics.py), quartic wall
// barrier (barrier.py), the warp-level rollout of Robot.forward / Robot.get_cost and the
// warp-level adjoint sweep of Robot.backward (klerg.py:409-450, 590-593, 686-710).
#pragma once
#include "klerg_common.cuh"

namespace klerg {

struct DynDev {
  int kind, S, A;
  float dt;
  int rpw[3];
  int has_map;
  float rot_lo[3], rot_hi[3], ang_lo[3], ang_hi[3];
};
struct BarDev {
  int n;
  float lo[KLERG_MAX_S], hi[KLERG_MAX_S], w[KLERG_MAX_S], pw[KLERG_MAX_S];
};

// Rarely-taken math is kept out of line: the fused kernels run these phases once per launch from a cold
// instruction cache, where straight-line code size is what costs time.
static __device__ __noinline__ float pow_generic(float d, float pw) { return powf(d, pw); }

__device__ __forceinline__ float powi_or_f(float d, float pw) {
  if (pw == 4.f) { const float d2 = d * d; return d2 * d2; }
  if (pw == 3.f) return d * d * d;
  if (pw == 2.f) return d * d;
  if (pw == 1.f) return d;
  return pow_generic(d, pw);
}

// barr(x) = sum_i w_i [(x_i <= lo_i)(x_i - lo_i)^pw + (x_i >= hi_i)(x_i - hi_i)^pw]   (barrier.py:70-76)
__device__ inline float barrier_value(const BarDev& b, const float* x) {
  float acc = 0.f;
  for (int i = 0; i < b.n; ++i) {
    const float xi = x[i];
    if (xi <= b.lo[i]) acc += b.w[i] * powi_or_f(xi - b.lo[i], b.pw[i]);
    if (xi >= b.hi[i]) acc += b.w[i] * powi_or_f(xi - b.hi[i], b.pw[i]);
  }
  return acc;
}

// dbarr (barrier.py:78-84); rows >= b.n stay zero
__device__ inline void barrier_grad(const BarDev& b, const float* x, int S, float* g) {
  for (int i = 0; i < S; ++i) {
    float acc = 0.f;
    if (i < b.n) {
      const float xi = x[i];
      if (xi <= b.lo[i]) acc += b.pw[i] * b.w[i] * powi_or_f(xi - b.lo[i], b.pw[i] - 1.f);
      if (xi >= b.hi[i]) acc += b.pw[i] * b.w[i] * powi_or_f(xi - b.hi[i], b.pw[i] - 1.f);
    }
    g[i] = acc;
  }
}

__device__ __forceinline__ float barrier_term(float x, float lo, float hi, float w, float pw) {
  float acc = 0.f;
  if (x <= lo) acc += w * powi_or_f(x - lo, pw);
  if (x >= hi) acc += w * powi_or_f(x - hi, pw);
  return acc;
}
__device__ __forceinline__ float barrier_dterm(float x, float lo, float hi, float w, float pw) {
  float acc = 0.f;
  if (x <= lo) acc += pw * w * powi_or_f(x - lo, pw - 1.f);
  if (x >= hi) acc += pw * w * powi_or_f(x - hi, pw - 1.f);
  return acc;
}

__device__ __forceinline__ float affine_map(float v, float ilo, float ihi, float olo, float ohi) {
  return (v - ilo) / (ihi - ilo) * (ohi - olo) + olo;
}

__device__ __forceinline__ float py_mod(float x, float m) {
  float r = fmodf(x, m);
  if (r < 0.f) r += m;
  return r;
}

static __device__ __noinline__ void sincos_ni(float x, float* s, float* c) { sincosf(x, s, c); }
static __device__ __noinline__ float atan2_ni(float y, float x) { return atan2f(y, x); }

// E = expm(hat(w) dt) via Rodrigues (= torch.matrix_exp of the skew matrix, dynamics.py:213-217).
// sin(th)/th and (1-cos(th))/th^2 from their series for th < 1 (truncation < 3e-8; |w| dt is a fraction of a
// radian for any admissible velocity), from sincos otherwise.
__device__ inline void rodrigues(const float* w, float dt, float* E) {
  const float kx = w[0] * dt, ky = w[1] * dt, kz = w[2] * dt;
  const float th2 = kx * kx + ky * ky + kz * kz;
  float A, B;
  if (th2 < 1.f) {
    A = 1.f + th2 * (-1.f / 6.f + th2 * (1.f / 120.f + th2 * (-1.f / 5040.f + th2 * (1.f / 362880.f))));
    B = 0.5f + th2 * (-1.f / 24.f + th2 * (1.f / 720.f + th2 * (-1.f / 40320.f + th2 * (1.f / 3628800.f))));
  } else {
    const float th = sqrtf(th2);
    float s, c;
    sincos_ni(th, &s, &c);
    A = s / th;
    B = (1.f - c) / th2;
  }
  // E = I + A K + B K^2,  K = hat(k)
  E[0] = 1.f - B * (ky * ky + kz * kz); E[1] = -A * kz + B * kx * ky;         E[2] = A * ky + B * kx * kz;
  E[3] = A * kz + B * kx * ky;          E[4] = 1.f - B * (kx * kx + kz * kz); E[5] = -A * kx + B * ky * kz;
  E[6] = -A * ky + B * kx * kz;         E[7] = A * kx + B * ky * kz;          E[8] = 1.f - B * (kx * kx + ky * ky);
}

// wrap(euler_XYZ(R)): roll in [0, 2pi), pitch / yaw in [-pi, pi)   (rotations.py:142-181, dynamics.py:219-222).
// atan2 / asin already return values within one period, so the reference's modulo reduces to one conditional
// add / subtract (same arithmetic as torch's `%` on that range).
__device__ inline void wrapped_euler_xyz(const float* Rn, float* rot) {
  const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
  float r0 = atan2_ni(Rn[7], Rn[8]);
  const float r1 = asinf(-Rn[6]);
  const float r2 = atan2_ni(Rn[3], Rn[0]);
  if (r0 < 0.f) r0 += two_pi;
  float t2 = r2 + pi;
  if (t2 >= two_pi) t2 -= two_pi;
  rot[0] = r0;
  rot[1] = (r1 + pi) - pi;
  rot[2] = t2 - pi;
}

__device__ inline void euler_xyz_to_matrix(const float* rot, float* R) {
  // Rz(yaw) * Ry(pitch) * Rx(roll)   (rotations.py:70-96, order flipped to match scipy)
  float sr, cr, sp, cp, sy, cy;
  sincos_ni(rot[0], &sr, &cr);
  sincos_ni(rot[1], &sp, &cp);
  sincos_ni(rot[2], &sy, &cy);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

// E(roll, pitch + 1e-5) R: the rpw x rpw block of d(pos rate)/d(vel)   (dynamics.py:189-211,283-289)
__device__ inline void euler_rate_block(const float* rot_in, const float* R, float* out9) {
  const float rot1 = rot_in[1] + 1e-5f;
  float s0, c0, s1, cc1;
  sincos_ni(rot_in[0], &s0, &c0);
  sincos_ni(rot1, &s1, &cc1);
  const float t1 = s1 / cc1;
  const float Em[9] = {1.f, s0 * t1, c0 * t1, 0.f, c0, -s0, 0.f, s0 / cc1, c0 / cc1};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out9[r * 3 + c] = Em[r * 3] * R[c] + Em[r * 3 + 1] * R[3 + c] + Em[r * 3 + 2] * R[6 + c];
}

__device__ __forceinline__ void matmul3(const float* E, const float* R, float* Rn) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = E[r * 3] * R[c] + E[r * 3 + 1] * R[3 + c] + E[r * 3 + 2] * R[6 + c];
}

// floats of scratch rollout_block needs at s_rot for the ROLL model: per candidate E_t [H][9], R_t [H+1][9]
__host__ __device__ inline int rollout_rot_floats(int G, int H) { return G * (2 * H + 1) * 9; }

// ---------------------------------------------------------------------------
// rollout of G control sequences by a whole CTA (every thread must call it).
//
// The RK4 step of the integrator models is exact and closed-form (A is nilpotent of index 2,
// dynamics.py:7-13,58-65), so only three short recurrences are serial: velocity / position
// running sums (one lane per candidate and control) and, for the ROLL model, R_{t+1} = E_t R_t
// (one lane per candidate).  Everything else - Rodrigues matrices E_t, Euler angles,
// linearisation blocks, wall barrier and its derivative - is evaluated for all time steps in
// parallel.
//
//   s_u     [G][H][A]    controls (shared)
//   s_traj  [G][H+1][S]  traj[t] = state before step t (Robot.forward) = state after step t-1 (get_cost)
//   s_dbarr [G][H][S]    dbarr(traj[t]), t < H            (may be NULL)
//   s_P     [G][H][A*A]  d(pos rate)/d(vel) block of A_t  (ROLL only; may be NULL)
//   s_rot   scratch of rollout_rot_floats(G, H) floats    (ROLL only)
//   s_red   scratch of 32 floats
//   s_bsum  [G]          sum_t barr(traj[t+1])  (klerg.py:708)
//   R_out   [G][9]       rotation after the last step (global or shared, may be NULL)
//   part    0 = everything; 1 = states only (running sums, rotation chain, angles): what the pair passes need;
//           2 = the rest (linearisation blocks P, dbarr, barrier sums) from the states and rotations that a
//               part-1 call left in s_traj / s_rot: only the CTA that runs the adjoint needs it.
// Ends with a __syncthreads().
// ---------------------------------------------------------------------------
// Running sum by ONE warp: out[k] = init + x[0] + ... + x[k] (inclusive) or init + x[0] + ... + x[k-1] (exclusive)
// for k < n, 32 elements per pass with a 5-step shuffle scan and a carry between passes.  `get(k)` supplies x[k],
// `put(k, v)` receives the result.  A lone warp executing a serial recurrence costs ~100 cycles per step (every
// instruction waits for its predecessor); the scan does 32 steps in about as many cycles.  Returns init + total.
template <typename Get, typename Put>
__device__ __forceinline__ float warp_running_sum(int n, float init, bool exclusive, Get get, Put put) {
  const int lane = threadIdx.x & 31;
  float carry = init;
  for (int base = 0; base < n; base += 32) {
    const int k = base + lane;
    const float own = k < n ? get(k) : 0.f;
    float v = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (k < n) put(k, carry + (exclusive ? v - own : v));
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
  return carry;
}

#ifdef KLERG_STAMPS
static __device__ long long g_ro_stamp[8];
#define RO_STAMP(i) if (blockIdx.x == 0 && threadIdx.x == 0) g_ro_stamp[i] = clock64()
#else
#define RO_STAMP(i)
#endif

__device__ inline void rollout_block(const DynDev& d, const BarDev& bar, const float* x0, const float* R0,
                                     const float* s_u, int G, int H, float* s_traj, float* s_dbarr, float* s_P,
                                     float* s_rot, float* s_red, float* s_bsum, float* R_out, int part = 0) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const PinnedParams pp = pin_params(s_red, d.S, d.A, H, d.kind, d.dt);
  const int S = pp.S, a = pp.A, kind = pp.kind;
  H = pp.H;
  const bool single = kind == KLERG_DYN_SINGLE, speed = kind == KLERG_DYN_SPEED, roll = kind == KLERG_DYN_ROLL;
  const float dt = pp.dt, c1 = 0.8f * dt, c2 = 0.4f * dt * dt;
  RO_STAMP(0);
  if (part != 2) {
  // (A) running sums, one warp per (candidate, control): vel_t = v0 + dt sum_{k<t} u_k, then
  //     pos_t = p0 + sum_{k<t} (0.8 dt vel_k + 0.4 dt^2 u_k)  (the closed-form RK4 step summed over time)
  for (int l = tid >> 5; l < G * a; l += nthr >> 5) {
    const int g = l / a, i = l - g * a;
    const float* us = s_u + (size_t)g * H * a;
    float* tr = s_traj + (size_t)g * (H + 1) * S;
    const float p0 = x0[i], v0 = single ? 0.f : x0[a + i];
    if ((tid & 31) == 0) {
      tr[i] = p0;
      if (!single) tr[a + i] = v0;
      if (speed) tr[2 * a + i] = x0[2 * a + i];
    }
    if (single) {
      warp_running_sum(H, p0, false, [&](int k) { return dt * us[k * a + i]; }, [&](int k, float v) { tr[(k + 1) * S + i] = v; });
    } else {
      warp_running_sum(H, v0, false, [&](int k) { return dt * us[k * a + i]; },
                       [&](int k, float v) {
                         tr[(k + 1) * S + a + i] = v;
                         if (speed) tr[(k + 1) * S + 2 * a + i] = fabsf(v);
                       });
      __syncwarp();
      warp_running_sum(H, p0, false, [&](int k) { return c1 * tr[k * S + a + i] + c2 * us[k * a + i]; },
                       [&](int k, float v) { tr[(k + 1) * S + i] = v; });
    }
  }
  __syncthreads();
  }
  RO_STAMP(1);
  if (roll) {
    float* s_E = s_rot;               // [G][H][9]
    float* s_R = s_rot + G * H * 9;   // [G][H+1][9]
    if (part != 2) {
    // (B) E_t from the pre-step angular velocity, all (g, t) in parallel; R_0
    for (int e = tid; e < G * H; e += nthr) {
      const int g = e / H, t = e - g * H;
      float w3[3], E[9];
      for (int k = 0; k < 3; ++k) w3[k] = s_traj[((size_t)g * (H + 1) + t) * S + a + d.rpw[k]];
      rodrigues(w3, dt, E);
      for (int i = 0; i < 9; ++i) s_E[e * 9 + i] = E[i];
    }
    if (tid == nthr - 1) {
      float R[9];
      if (R0) {
        for (int i = 0; i < 9; ++i) R[i] = R0[i];
      } else {
        float rot[3];
        for (int k = 0; k < 3; ++k) {
          rot[k] = x0[d.rpw[k]];
          if (d.has_map) rot[k] = affine_map(rot[k], d.rot_lo[k], d.rot_hi[k], d.ang_lo[k], d.ang_hi[k]);
        }
        euler_xyz_to_matrix(rot, R);
      }
      for (int g = 0; g < G; ++g)
        for (int i = 0; i < 9; ++i) s_R[(size_t)g * (H + 1) * 9 + i] = R[i];
    }
    __syncthreads();
    RO_STAMP(5);
    // (C) R_{t+1} = E_t E_{t-1} ... E_0 R_0: matrix products are associative, so one warp per candidate forms
    //     the running products with a shuffle scan (32 steps per pass) instead of a serial chain
    for (int g = tid >> 5; g < G; g += nthr >> 5) {
      const int lane = tid & 31;
      float* Rg = s_R + (size_t)g * (H + 1) * 9;
      const float* Eg = s_E + (size_t)g * H * 9;
      float carry[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) carry[i] = Rg[i];
      for (int base = 0; base < H; base += 32) {
        const int k = base + lane;
        float v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = (k < H) ? Eg[k * 9 + i] : ((i % 4 == 0) ? 1.f : 0.f);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float t[9], n[9];
#pragma unroll
          for (int i = 0; i < 9; ++i) t[i] = __shfl_up_sync(0xffffffffu, v[i], o);
          matmul3(v, t, n);  // later * earlier
          if (lane >= o) {
#pragma unroll
            for (int i = 0; i < 9; ++i) v[i] = n[i];
          }
        }
        float r[9];
        matmul3(v, carry, r);
        if (k < H) {
#pragma unroll
          for (int i = 0; i < 9; ++i) Rg[(k + 1) * 9 + i] = r[i];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) carry[i] = __shfl_sync(0xffffffffu, r[i], 31);
      }
    }
    __syncthreads();
    RO_STAMP(6);
    // (D1) angles of R_t overwrite the integrated angle positions (dynamics.py:291-301)
    for (int e = tid; e < G * H; e += nthr) {
      const int g = e / H, t = 1 + (e - g * H);
      float rot[3];
      wrapped_euler_xyz(s_R + ((size_t)g * (H + 1) + t) * 9, rot);
      for (int k = 0; k < 3; ++k) {
        float v = rot[k];
        if (d.has_map) v = affine_map(v, d.ang_lo[k], d.ang_hi[k], d.rot_lo[k], d.rot_hi[k]);
        s_traj[((size_t)g * (H + 1) + t) * S + d.rpw[k]] = v;
      }
    }
    if (R_out)
      for (int e = tid; e < G * 9; e += nthr) R_out[e] = s_R[((size_t)(e / 9) * (H + 1) + H) * 9 + e % 9];
    __syncthreads();
    }
    RO_STAMP(7);
    // (D2) linearisation block: 0.8 I with the rpw x rpw entries replaced by E(rot) R (dynamics.py:189-211,283-289)
    if (s_P && part != 1) {
      for (int e = tid; e < G * H * a * a; e += nthr) {
        const int r = (e % (a * a)) / a, c = e % a;
        s_P[e] = (r == c) ? 0.8f : 0.f;
      }
      __syncthreads();
      for (int e = tid; e < G * H; e += nthr) {
        const int g = e / H, t = e - g * H;
        float rot[3];
        for (int k = 0; k < 3; ++k) {
          rot[k] = s_traj[((size_t)g * (H + 1) + t) * S + d.rpw[k]];
          if (d.has_map) rot[k] = affine_map(rot[k], d.rot_lo[k], d.rot_hi[k], d.ang_lo[k], d.ang_hi[k]);
        }
        float blk[9];
        euler_rate_block(rot, s_R + ((size_t)g * (H + 1) + t) * 9, blk);
        float* Pt = s_P + (size_t)e * a * a;
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) Pt[d.rpw[r] * a + d.rpw[c]] = blk[r * 3 + c];
      }
    }
  } else if (R_out && part != 2) {
    for (int e = tid; e < G * 9; e += nthr) R_out[e] = ((e % 9) % 4 == 0) ? 1.f : 0.f;
  }
  RO_STAMP(2);
  if (part == 1) {
    __syncthreads();
    return;
  }
  // (E) wall barrier of the post-step states and its derivative at the pre-step states
  for (int g = 0; g < G; ++g) {
    const float* tr = s_traj + (size_t)g * (H + 1) * S;
    float bsum = 0.f;
    for (int e = tid; e < H * S; e += nthr) {
      const int t = e / S, i = e - t * S;
      float db = 0.f;
      if (i < bar.n) {
        const float lo = bar.lo[i], hi = bar.hi[i], w = bar.w[i], pw = bar.pw[i];
        bsum += barrier_term(tr[(t + 1) * S + i], lo, hi, w, pw);
        if (s_dbarr) db = barrier_dterm(tr[e], lo, hi, w, pw);
      }
      if (s_dbarr) s_dbarr[(size_t)g * H * S + e] = db;
    }
    RO_STAMP(3);
    bsum = warp_sum_f(bsum);
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = bsum;
    __syncthreads();
    if (tid == 0) {
      float v = 0.f;
      for (int w = 0; w < (nthr + 31) / 32; ++w) v += s_red[w];
      s_bsum[g] = v;
    }
  }
  __syncthreads();
  RO_STAMP(4);
}

// ---------------------------------------------------------------------------
// adjoint sweep (klerg.py:433-450, 590-593), default policy (dmudx = 0), by a whole CTA.
// rho_H = 0; for t = H-1..0 one RK4 step of rho' = g_t - A_t^T rho with step -dt, g = dgdx - dbarr:
//   rho_p <- rho_p + h g_p,   rho_v <- rho_v + h (g_v - P^T rho_p) - h^2/2 P^T g_p        (h = -dt)
// rho_p does not depend on rho_v, so the sweep is three running sums (one lane per control) around
// a time-parallel P^T product; du_t = -Rinv B^T rho, djdlam_t = rho B du_t, u* = clamp(u + alpha du)
// are evaluated for all t in parallel.
//   sg [H][S] (shared), sP [H][A*A] or NULL (0.8 I), s_traj for sign(vel) (SPEED), su [H][A],
//   s_scr scratch of adjoint_scratch_floats(H, A) floats.   Ends with a __syncthreads().
// ---------------------------------------------------------------------------
struct AdjParams {
  float rinv[KLERG_MAX_A];
  float clo[KLERG_MAX_A], chi[KLERG_MAX_A];
  float alpha;
  float sat;  // > 0: u* = tanh(us / sat) * chi (Robot.saturate_control, klerg.py:342-349) instead of the clamp
};

__host__ __device__ inline int adjoint_scratch_floats(int H, int A) { return 5 * H * A + 8; }

__device__ inline void adjoint_block(const DynDev& d, const AdjParams& ap, int H, const float* sg, const float* sP,
                                     const float* s_traj, const float* su, float* s_scr, float* du, float* djdlam,
                                     float* u_star) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const PinnedParams pp = pin_params(s_scr + 5 * H * d.A, d.S, d.A, H, d.kind, d.dt);
  const int S = pp.S, A = pp.A, kind = pp.kind;
  H = pp.H;
  const bool single = kind == KLERG_DYN_SINGLE, speed = kind == KLERG_DYN_SPEED;
  const float h = -pp.dt;
  float* s_rp = s_scr;              // rho_p before the update of step t
  float* s_rm = s_rp + H * A;       // rho_m after the update (SPEED)
  float* s_a = s_rm + H * A;        // g_v - P^T rho_p
  float* s_b = s_a + H * A;         // P^T g_p
  float* s_btr = s_b + H * A;       // B^T rho after step t
  // rho_p (and rho_m) are running sums of h*g from the end of the horizon: one warp per control
  for (int i = tid >> 5; i < A; i += nthr >> 5) {
    // value BEFORE the update of step t = sum over later steps (exclusive), value after = inclusive
    warp_running_sum(H, 0.f, !single, [&](int k) { return h * sg[(H - 1 - k) * S + i]; },
                     [&](int k, float v) {
                       if (single) s_btr[(H - 1 - k) * A + i] = v;
                       else s_rp[(H - 1 - k) * A + i] = v;
                     });
    if (speed)
      warp_running_sum(H, 0.f, false, [&](int k) { return h * sg[(H - 1 - k) * S + 2 * A + i]; },
                       [&](int k, float v) { s_rm[(H - 1 - k) * A + i] = v; });
  }
  __syncthreads();
  if (!single) {
    for (int e = tid; e < H * A; e += nthr) {
      const int t = e / A, i = e - t * A;
      float ptr_ = 0.f, ptg = 0.f;
      if (sP) {
        const float* Pt = sP + t * A * A;
        for (int k = 0; k < A; ++k) {
          const float pk = Pt[k * A + i];
          ptr_ = fmaf(pk, s_rp[t * A + k], ptr_);
          ptg = fmaf(pk, sg[t * S + k], ptg);
        }
      } else {
        ptr_ = 0.8f * s_rp[e];
        ptg = 0.8f * sg[t * S + i];
      }
      s_a[e] = sg[t * S + A + i] - ptr_;
      s_b[e] = ptg;
    }
    __syncthreads();
    for (int i = tid >> 5; i < A; i += nthr >> 5) {
      warp_running_sum(H, 0.f, false,
                       [&](int k) {
                         const int t = H - 1 - k;
                         return h * s_a[t * A + i] - 0.5f * h * h * s_b[t * A + i];
                       },
                       [&](int k, float rv) {
                         const int t = H - 1 - k;
                         float btr = rv;
                         if (speed) btr = rv + ((s_traj[t * S + A + i] < 0.f) ? -1.f : 1.f) * s_rm[t * A + i];
                         s_btr[t * A + i] = btr;
                       });
    }
    __syncthreads();
  }
  for (int t = tid; t < H; t += nthr) {
    float dj = 0.f;
    for (int i = 0; i < A; ++i) {
      const float btr = s_btr[t * A + i];
      const float dui = -ap.rinv[i] * btr;
      dj += btr * dui;
      du[t * A + i] = dui;
      const float us = su[t * A + i] + ap.alpha * dui;
      u_star[t * A + i] = ap.sat > 0.f ? tanhf(us / ap.sat) * ap.chi[i] : fminf(fmaxf(us, ap.clo[i]), ap.chi[i]);
    }
    djdlam[t] = dj;
  }
  __syncthreads();
}

bool make_dyn(const klerg_dyn_spec* s, DynDev& d);
bool make_bar(const klerg_barrier_spec* s, BarDev& b);

}  // namespace klerg

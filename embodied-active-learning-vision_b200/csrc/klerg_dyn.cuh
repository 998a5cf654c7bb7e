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

__device__ __forceinline__ float powi_or_f(float d, float pw) {
  if (pw == 4.f) { const float d2 = d * d; return d2 * d2; }
  if (pw == 3.f) return d * d * d;
  if (pw == 2.f) return d * d;
  if (pw == 1.f) return d;
  return powf(d, pw);
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

__device__ inline void euler_xyz_to_matrix(const float* rot, float* R) {
  // Rz(yaw) * Ry(pitch) * Rx(roll)   (rotations.py:70-96, order flipped to match scipy)
  float sr, cr, sp, cp, sy, cy;
  sincosf(rot[0], &sr, &cr);
  sincosf(rot[1], &sp, &cp);
  sincosf(rot[2], &sy, &cy);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

__device__ __forceinline__ float py_mod(float x, float m) {
  float r = fmodf(x, m);
  if (r < 0.f) r += m;
  return r;
}

// Rn = expm(hat(w) dt) * R via Rodrigues; new angles = wrap(euler_XYZ(Rn))   (dynamics.py:213-222)
__device__ inline void advance_rotation(const float* R, const float* w, float dt, float* Rn, float* rot) {
  const float kx = w[0] * dt, ky = w[1] * dt, kz = w[2] * dt;
  const float th2 = kx * kx + ky * ky + kz * kz;
  float A, B;  // sin(th)/th, (1-cos(th))/th^2
  if (th2 < 1e-8f) {
    A = 1.f - th2 / 6.f;
    B = 0.5f - th2 / 24.f;
  } else {
    const float th = sqrtf(th2);
    float s, c;
    sincosf(th, &s, &c);
    A = s / th;
    B = (1.f - c) / th2;
  }
  // E = I + A K + B K^2,  K = hat(k)
  float E[9];
  E[0] = 1.f - B * (ky * ky + kz * kz); E[1] = -A * kz + B * kx * ky;         E[2] = A * ky + B * kx * kz;
  E[3] = A * kz + B * kx * ky;          E[4] = 1.f - B * (kx * kx + kz * kz); E[5] = -A * kx + B * ky * kz;
  E[6] = -A * ky + B * kx * kz;         E[7] = A * kx + B * ky * kz;          E[8] = 1.f - B * (kx * kx + ky * ky);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = E[r * 3] * R[c] + E[r * 3 + 1] * R[3 + c] + E[r * 3 + 2] * R[6 + c];
  const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
  float r0 = atan2f(Rn[7], Rn[8]);
  float r1 = asinf(-Rn[6]);
  float r2 = atan2f(Rn[3], Rn[0]);
  rot[0] = py_mod(r0, two_pi);
  rot[1] = py_mod(r1 + pi, two_pi) - pi;
  rot[2] = py_mod(r2 + pi, two_pi) - pi;
}

// ---------------------------------------------------------------------------
// rollout: ONE WARP rolls one control sequence us[H][A] out from x0 with RK4 (closed form:
// A is nilpotent of index 2, so RK4 == the exact cubic; dynamics.py:7-13,58-65).
// Lane i < A owns (pos_i, vel_i[, mag_i]); every lane carries the ROLL rotation matrix.
//   traj[t]  (t = 0..H)   state before step t / after step t-1, [H+1][S]    (may be NULL)
//   dbarr[t] (t < H)      dbarr(traj[t]), [H][S]                            (may be NULL)
//   P[t]     (t < H)      d(pos rate)/d(vel) block of A_t, [H][A*A]         (may be NULL)
// Pointers may be global or shared.  Returns sum_t barr(traj[t+1]) (klerg.py:708) in every lane.
// ---------------------------------------------------------------------------
__device__ inline float rollout_warp(const DynDev& d, const BarDev& bar, const float* x0, const float* R0,
                                     const float* us, int H, float* traj, float* dbarr, float* P, float* R_out) {
  const int lane = threadIdx.x & 31;
  const int S = d.S, a = d.A;
  const bool single = d.kind == KLERG_DYN_SINGLE, speed = d.kind == KLERG_DYN_SPEED, roll = d.kind == KLERG_DYN_ROLL;
  const bool act = lane < a;
  float pos = act ? x0[lane] : 0.f;
  float vel = (act && !single) ? x0[a + lane] : 0.f;
  float mag = (act && speed) ? x0[2 * a + lane] : 0.f;
  // barrier rows owned by this lane: position row `lane`, velocity row `a + lane`, magnitude row `2a + lane`
  float blo_p = 0.f, bhi_p = 0.f, bw_p = 0.f, bpw_p = 1.f, blo_v = 0.f, bhi_v = 0.f, bw_v = 0.f, bpw_v = 1.f;
  float blo_m = 0.f, bhi_m = 0.f, bw_m = 0.f, bpw_m = 1.f;
  bool has_p = false, has_v = false, has_m = false;
  if (act && lane < bar.n) { has_p = true; blo_p = bar.lo[lane]; bhi_p = bar.hi[lane]; bw_p = bar.w[lane]; bpw_p = bar.pw[lane]; }
  if (act && !single && a + lane < bar.n) { has_v = true; blo_v = bar.lo[a + lane]; bhi_v = bar.hi[a + lane]; bw_v = bar.w[a + lane]; bpw_v = bar.pw[a + lane]; }
  if (act && speed && 2 * a + lane < bar.n) { has_m = true; blo_m = bar.lo[2 * a + lane]; bhi_m = bar.hi[2 * a + lane]; bw_m = bar.w[2 * a + lane]; bpw_m = bar.pw[2 * a + lane]; }

  float R[9];
  int my_rot = -1;  // which of roll/pitch/yaw this lane's position is (ROLL)
  if (roll) {
    for (int k = 0; k < 3; ++k)
      if (lane == d.rpw[k]) my_rot = k;
    float rot[3];
    for (int k = 0; k < 3; ++k) {
      rot[k] = __shfl_sync(0xffffffffu, pos, d.rpw[k]);
      if (d.has_map) rot[k] = affine_map(rot[k], d.rot_lo[k], d.rot_hi[k], d.ang_lo[k], d.ang_hi[k]);
    }
    if (R0) {
      for (int i = 0; i < 9; ++i) R[i] = R0[i];
    } else {
      euler_xyz_to_matrix(rot, R);
    }
  }
  const float dt = d.dt, c1 = 0.8f * dt, c2 = 0.4f * dt * dt;
  float bsum = 0.f;
  for (int t = 0; t <= H; ++t) {
    if (act && traj) {
      traj[t * S + lane] = pos;
      if (!single) traj[t * S + a + lane] = vel;
      if (speed) traj[t * S + 2 * a + lane] = mag;
    }
    if (t > 0 && act) {
      if (has_p) bsum += barrier_term(pos, blo_p, bhi_p, bw_p, bpw_p);
      if (has_v) bsum += barrier_term(vel, blo_v, bhi_v, bw_v, bpw_v);
      if (has_m) bsum += barrier_term(mag, blo_m, bhi_m, bw_m, bpw_m);
    }
    if (t == H) break;
    if (dbarr && act) {
      float* db = dbarr + t * S;
      db[lane] = has_p ? barrier_dterm(pos, blo_p, bhi_p, bw_p, bpw_p) : 0.f;
      if (!single) db[a + lane] = has_v ? barrier_dterm(vel, blo_v, bhi_v, bw_v, bpw_v) : 0.f;
      if (speed) db[2 * a + lane] = has_m ? barrier_dterm(mag, blo_m, bhi_m, bw_m, bpw_m) : 0.f;
    }
    float w3[3] = {0.f, 0.f, 0.f}, rot3[3] = {0.f, 0.f, 0.f};
    if (roll) {
      for (int k = 0; k < 3; ++k) {
        w3[k] = __shfl_sync(0xffffffffu, vel, d.rpw[k]);
        rot3[k] = __shfl_sync(0xffffffffu, pos, d.rpw[k]);
      }
    }
    if (P) {
      // 0.8 I, with the rpw x rpw entries replaced by E(rot) R   (dynamics.py:189-211,283-289)
      float* Pt = P + t * a * a;
      for (int e = lane; e < a * a; e += 32) Pt[e] = (e / a == e % a) ? 0.8f : 0.f;
      __syncwarp();
      if (roll && lane == 0) {
        float rot[3];
        for (int k = 0; k < 3; ++k)
          rot[k] = d.has_map ? affine_map(rot3[k], d.rot_lo[k], d.rot_hi[k], d.ang_lo[k], d.ang_hi[k]) : rot3[k];
        rot[1] += 1e-5f;
        float s0, c0;
        sincosf(rot[0], &s0, &c0);
        const float t1 = tanf(rot[1]), cc1 = cosf(rot[1]);
        const float Em[9] = {1.f, s0 * t1, c0 * t1, 0.f, c0, -s0, 0.f, s0 / cc1, c0 / cc1};
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c)
            Pt[d.rpw[r] * a + d.rpw[c]] = Em[r * 3] * R[c] + Em[r * 3 + 1] * R[3 + c] + Em[r * 3 + 2] * R[6 + c];
      }
    }
    const float ut = act ? us[t * a + lane] : 0.f;
    if (single) {
      pos = pos + dt * ut;
    } else {
      pos = pos + (c1 * vel + c2 * ut);
      vel = vel + dt * ut;
      if (speed) mag = fabsf(vel);
    }
    if (roll) {
      float Rn[9], nr[3];
      advance_rotation(R, w3, dt, Rn, nr);  // every lane computes it redundantly (no divergence)
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = Rn[i];
      if (my_rot >= 0) {
        float v = my_rot == 0 ? nr[0] : (my_rot == 1 ? nr[1] : nr[2]);
        if (d.has_map) v = affine_map(v, d.ang_lo[my_rot], d.ang_hi[my_rot], d.rot_lo[my_rot], d.rot_hi[my_rot]);
        pos = v;
      }
    }
  }
  bsum = warp_sum_f(bsum);
  if (lane == 0 && R_out) {
    if (!roll) {
      for (int i = 0; i < 9; ++i) R_out[i] = (i % 4 == 0) ? 1.f : 0.f;
    } else {
      for (int i = 0; i < 9; ++i) R_out[i] = R[i];
    }
  }
  return bsum;
}

// ---------------------------------------------------------------------------
// adjoint sweep (klerg.py:433-450, 590-593), default policy (dmudx = 0): ONE WARP.
// rho_H = 0; for t = H-1..0 one RK4 step of rho' = g_t - A_t^T rho with step -dt, where
// g = dgdx - dbarr is staged in shared memory; du_t = -Rinv B^T rho, djdlam_t = rho B du_t,
// u_star = clamp(u + alpha du).  Lane i < A carries component i of rho_p, rho_v (rho_m for SPEED).
//   sg [H][S] (smem), sP [H][A*A] or NULL (0.8 I), ssgn [H][A] sign(vel) (SPEED), su [H][A].
// ---------------------------------------------------------------------------
struct AdjParams {
  float rinv[KLERG_MAX_A];
  float clo[KLERG_MAX_A], chi[KLERG_MAX_A];
  float alpha;
};

__device__ inline void adjoint_warp(const DynDev& d, const AdjParams& ap, int H, const float* sg, const float* sP,
                                    const float* ssgn, const float* su, float* du, float* djdlam, float* u_star) {
  const int S = d.S, A = d.A;
  const bool single = d.kind == KLERG_DYN_SINGLE, speed = d.kind == KLERG_DYN_SPEED;
  const int i = threadIdx.x & 31;
  const bool act = i < A;
  const float h = -d.dt;
  const float rinv = act ? ap.rinv[i] : 0.f, clo = act ? ap.clo[i] : 0.f, chi = act ? ap.chi[i] : 0.f;
  float rp = 0.f, rv = 0.f, rm = 0.f;
  for (int t = H - 1; t >= 0; --t) {
    float gp = 0.f, gv = 0.f, gm = 0.f;
    if (act) {
      gp = sg[t * S + i];
      if (!single) gv = sg[t * S + A + i];
      if (speed) gm = sg[t * S + 2 * A + i];
    }
    float btr;  // (B^T rho)_i
    if (single) {
      rp = rp + h * gp;
      btr = rp;
    } else {
      float ptr_ = 0.f, ptg = 0.f;  // (P^T rho_p)_i, (P^T g_p)_i
      if (sP) {
        const float* Pt = sP + t * A * A;
        for (int kk = 0; kk < A; ++kk) {
          const float rk = __shfl_sync(0xffffffffu, rp, kk);
          const float gk = __shfl_sync(0xffffffffu, gp, kk);
          const float pk = act ? Pt[kk * A + i] : 0.f;
          ptr_ = fmaf(pk, rk, ptr_);
          ptg = fmaf(pk, gk, ptg);
        }
      } else {
        ptr_ = 0.8f * rp;
        ptg = 0.8f * gp;
      }
      const float rv_n = rv + h * (gv - ptr_) - 0.5f * h * h * ptg;
      rp = rp + h * gp;
      rv = rv_n;
      btr = rv;
      if (speed) {
        rm = rm + h * gm;
        btr = rv + (act ? ssgn[t * A + i] : 0.f) * rm;
      }
    }
    const float dui = act ? -rinv * btr : 0.f;
    const float dj = warp_sum_f(act ? btr * dui : 0.f);
    if (act) {
      du[t * A + i] = dui;
      const float us = su[t * A + i] + ap.alpha * dui;
      u_star[t * A + i] = fminf(fmaxf(us, clo), chi);
    }
    if (i == 0) djdlam[t] = dj;
  }
}

bool make_dyn(const klerg_dyn_spec* s, DynDev& d);
bool make_bar(const klerg_barrier_spec* s, BarDev& b);

}  // namespace klerg

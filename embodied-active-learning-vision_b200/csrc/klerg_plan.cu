// The optimisation loop of Robot.kldiv_planner (klerg.py:505-576) and its line search (klerg.py:712-751) without the
// host in it (sm_100a).
//
// The reference - and this library's host mirror when plot data is kept - reads djdlam / u* back after every gradient
// eval, takes the argmin, builds the <= 5 line-search candidates, reads their costs back, applies the accept rule and
// goes round again: two device-to-host round trips per iteration, which is most of a Robot.step() at the reference's
// own operating sizes (1e3 .. 1e5 samples).  Here the whole loop is ENQUEUED at once:
//
//   cost(u) -> begin | { gradient(u) -> select -> costs(candidates) -> accept } x num_iters | finish -> rollout(u)
//
// `select` / `accept` / `finish` are one-CTA kernels that keep the loop state (last cost, current plan, "still
// running") in a small device struct; the fused evals are the same launches the host loop makes, gated by that struct
// (EvalArgs::gate): once the loop has stopped - the reference's `break` - the remaining evals return immediately, and
// a cost launch evaluates exactly the candidates the host loop would have passed.  Decisions use the same float32
// comparisons as the host code, in the same order, so the plan is the one the host loop produces; with several ranks
// every rank takes the same decisions from the same all-reduced numbers.  The host reads ONE packed buffer back.
#include <cfloat>

#include "klerg_fused.cuh"

extern "C" int klerg_eval_costs(const klerg_kernel_spec*, const klerg_dyn_spec*, const klerg_barrier_spec*, const klerg_peers*,
                                const float*, const float*, const float*, int64_t, int64_t, const float*, int64_t, int64_t,
                                const float*, const float*, const double*, float, float*, float*, double*, float*, float*,
                                void*, void*);
extern "C" int klerg_eval_gradient(const klerg_kernel_spec*, const klerg_dyn_spec*, const klerg_barrier_spec*, const klerg_peers*,
                                   const float*, const float*, const float*, int64_t, const float*, int64_t, int64_t,
                                   const float*, const float*, const double*, float, const float*, float, const float*,
                                   const float*, float*, float*, double*, float*, float*, float*, float*, float*, double*,
                                   float*, void*, void*);
extern "C" int klerg_rollout(const klerg_dyn_spec*, const klerg_barrier_spec*, const float*, const float*, const float*, int64_t,
                             int64_t, float*, float*, float*, float*, float*, void*);

namespace klerg {

constexpr int PLAN_MAXC = 6;  // line-search windows (<= max_app_dur = 5) + the fallback window of a negative cost

struct PlanDev {
  float last_cost;
  int active;          // the loop is still running: gate of the gradient evals
  int n_cand;          // candidates of the pending cost eval: gate of the cost evals
  int t_app, lam0, n_win, extra;
  int win[PLAN_MAXC][2];
  int cost_evals, grad_evals, iters;
};

__global__ void plan_begin_kernel(PlanDev* st, const float* cost) {
  if (threadIdx.x == 0) {
    st->last_cost = cost[0];
    st->active = 1;
    st->n_cand = 0;
    st->cost_evals = 1;
    st->grad_evals = 0;
    st->iters = 0;
  }
}

// a Python slice bound on a sequence of length H
__device__ __forceinline__ int slice_bound(int i, int H) { return i < 0 ? max(i + H, 0) : min(i, H); }

// After a gradient eval: t_app = argmin djdlam (torch.argmin: the first minimum, a NaN wins), the windows the line
// search would try (klerg.py:714-738) and one candidate control sequence per window.
__global__ void plan_select_kernel(PlanDev* st, int idx, int H, int A, const float* djdlam, const float* u_star, const float* u,
                                   float* cands, int fixed_lam, int lam_fixed, int max_app_dur) {
  __shared__ int s_n;
  if (threadIdx.x == 0) {
    s_n = 0;
    st->n_cand = 0;
    if (st->active) {
      st->grad_evals += 1;
      int t = 0;
      float best = djdlam[0];
      for (int i = 1; i < H && best == best; ++i) {
        const float v = djdlam[i];
        if (v != v || v < best) {
          best = v;
          t = i;
        }
      }
      if (!(best < 0.f)) {
        st->active = 0;  // klerg.py:563-565: no descent direction, keep the plan
      } else {
        st->t_app = t;
        int n = 0;
        st->extra = -1;
        if (fixed_lam) {
          st->win[0][0] = t;
          st->win[0][1] = t + lam_fixed;
          n = 1;
          st->n_win = 1;
        } else {
          int lam;
          if (t == 0 || t == H - 1) lam = min(H, max_app_dur);
          else if (t == idx) lam = min(H - t, max_app_dur);
          else lam = min(min(t - idx, H - t - idx), (max_app_dur + 1) / 2);
          lam = max(lam, 1);
          st->lam0 = lam;
          for (; lam > 0 && n < PLAN_MAXC - 1; --lam, ++n) {
            if (t == idx) { st->win[n][0] = t; st->win[n][1] = lam + 1; }
            else if (t == H - 1) { st->win[n][0] = lam - 1; st->win[n][1] = t; }
            else { st->win[n][0] = t - lam; st->win[n][1] = t + lam + 1; }
          }
          st->n_win = n;
          if (st->last_cost < 0.f) {  // the select rule can stop before the first window: its window is [idx, lam0)
            st->extra = n;
            st->win[n][0] = idx;
            st->win[n][1] = st->lam0;
            ++n;
          }
        }
        st->n_cand = n;
        s_n = n;
      }
    }
  }
  __syncthreads();
  const int n = s_n;
  if (n == 0) return;
  const int t_app = st->t_app;
  for (int e = threadIdx.x; e < n * H * A; e += blockDim.x) {
    const int c = e / (H * A), r = e - c * H * A, t = r / A, a = r - t * A;
    const int lo = slice_bound(st->win[c][0], H), hi = slice_bound(st->win[c][1], H);
    cands[e] = (t >= lo && t < hi) ? u_star[t_app * A + a] : u[r];
  }
}

// After the candidates' costs: the sequential accept / stop rule of the line search (klerg.py:723-751), the planner's
// own accept rule (klerg.py:566-574), and the new plan.
__global__ void plan_accept_kernel(PlanDev* st, int idx, int H, int A, const float* Js, const float* cands, float* u, int fixed_lam) {
  __shared__ int s_take;
  if (threadIdx.x == 0) {
    s_take = -1;
    if (st->active) {
      const float J0 = st->last_cost;
      float cost = J0;
      int take = -1;  // candidate that becomes the plan (-1: the plan stays)
      if (fixed_lam) {
        cost = Js[0];
        take = 0;
        st->cost_evals += 1;
      } else {
        const int nw = st->n_win;
        st->cost_evals += nw;
        float Jn = J0 * 2.f;
        bool done = false;
        int k = 0, last_k = -1, cur = -1;  // cur: window of Jn; last_k: window chosen when the rule stops
        bool last_is_init = true;
        while (!done && k < nw) {
          const float Jn_last = Jn;
          const int prev = cur;
          Jn = Js[k];
          cur = k;
          ++k;
          if (Jn_last < J0 && Jn > Jn_last) {
            done = true;
            last_k = prev;
            last_is_init = prev < 0;
          }
        }
        if (!done && Jn < J0) {
          take = cur;
          cost = Js[cur];
        } else if (done) {
          if (last_is_init) {  // stopped before the first window: the initial window [idx, lam0), evaluated as `extra`
            take = st->extra;
            cost = Js[take];
            st->cost_evals += 1;
          } else {
            take = last_k;
            cost = Js[last_k];
          }
        }
      }
      if (idx > 0 && J0 <= cost) {
        st->active = 0;  // klerg.py:569-571: no improvement, keep the previous iterate
      } else {
        st->last_cost = cost;
        st->iters += 1;
        s_take = take;
      }
    }
    st->n_cand = 0;
  }
  __syncthreads();
  const int take = s_take;
  if (take < 0) return;
  for (int e = threadIdx.x; e < H * A; e += blockDim.x) u[e] = cands[(size_t)take * H * A + e];
}

// nan_to_num of the plan (klerg.py:576) and the packed results: {last_cost, fault, cost evals, gradient evals,
// accepted iterations, 0, 0, 0, u[H][A]}; the final rollout is written behind it by klerg_rollout.
__global__ void plan_finish_kernel(const PlanDev* st, int HA, float* u, float* out, const unsigned* ctrl) {
  if (threadIdx.x == 0) {
    out[0] = st->last_cost;
    out[1] = ctrl[5] ? 1.f : 0.f;
    out[2] = (float)st->cost_evals;
    out[3] = (float)st->grad_evals;
    out[4] = (float)st->iters;
    out[5] = out[6] = out[7] = 0.f;
  }
  for (int e = threadIdx.x; e < HA; e += blockDim.x) {
    float v = u[e];
    if (v != v) v = 0.f;
    else if (v > FLT_MAX) v = FLT_MAX;
    else if (v < -FLT_MAX) v = -FLT_MAX;
    u[e] = v;
    out[8 + e] = v;
  }
}

struct PlanScratch {
  PlanDev* st;
  float *cands, *cost_pack, *djdlam, *u_star, *du, *dgdx, *v_grad, *v_costs;
  double* totals;
  size_t bytes;
};

static PlanScratch plan_layout(void* base, int64_t H, int S, int A, int64_t ld) {
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return (char*)base + at; };
  PlanScratch s{};
  s.st = (PlanDev*)take(sizeof(PlanDev));
  s.cands = (float*)take(sizeof(float) * PLAN_MAXC * H * A);
  s.cost_pack = (float*)take(sizeof(float) * (FUSED_MAXG + 1));
  s.djdlam = (float*)take(sizeof(float) * H);
  s.u_star = (float*)take(sizeof(float) * H * A);
  s.du = (float*)take(sizeof(float) * H * A);
  s.dgdx = (float*)take(sizeof(float) * H * S);
  s.totals = (double*)take(sizeof(double) * 2 * FUSED_MAXG);
  s.v_grad = (float*)take(sizeof(float) * ld);
  s.v_costs = (float*)take(sizeof(float) * PLAN_MAXC * ld);
  s.bytes = o;
  return s;
}

}  // namespace klerg

using namespace klerg;

extern "C" size_t klerg_plan_scratch_bytes(int64_t H, int32_t S, int32_t A, int64_t ld) {
  return plan_layout(nullptr, H, S, A, ld).bytes;
}

extern "C" int64_t klerg_plan_result_floats(int64_t H, int32_t S, int32_t A) { return 8 + H * A + (H + 1) * S; }

extern "C" int klerg_plan_optimize(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                                   const klerg_peers* peers, const float* x0, const float* R0, float* u, int64_t H,
                                   const float* packed, int64_t N, int64_t ld, const float* q_base, const float* p,
                                   const double* p_stats, float floor, const float* Rinv_diag, float alpha,
                                   const float* ctrl_lo, const float* ctrl_hi, int32_t num_iters, int32_t fixed_lam,
                                   int32_t lam, int32_t max_app_dur, void* scratch, float* result, void* workspace,
                                   void* stream) {
  if (!dyn || !u || !scratch || !result || !workspace) { set_error("plan_optimize: null argument"); return -1; }
  if (H < 1 || H > KLERG_MAX_H) { set_error("plan_optimize: H out of range"); return -1; }
  if (num_iters < 1 || num_iters > 4096) { set_error("plan_optimize: num_iters out of range"); return -1; }
  if (max_app_dur < 1 || max_app_dur > PLAN_MAXC - 1) { set_error("plan_optimize: max_app_dur must be in 1..%d", PLAN_MAXC - 1); return -1; }
  if (fixed_lam && lam < 1) { set_error("plan_optimize: lam must be positive"); return -1; }
  if (ld & 3) { set_error("plan_optimize: bad sample stride"); return -1; }
  const int S = dyn->S, A = dyn->A;
  const PlanScratch s = plan_layout(scratch, H, S, A, ld);
  cudaStream_t st = (cudaStream_t)stream;
  const int ngate = fixed_lam ? 1 : PLAN_MAXC;
  int rc = 0;
  struct GateReset { ~GateReset() { g_eval_gate = nullptr; } } reset;  // never leave the gate set behind an early return

  rc = klerg_eval_costs(k, dyn, bar, peers, x0, R0, u, 1, H, packed, N, ld, q_base, p, p_stats, floor, s.v_costs, nullptr,
                        nullptr, s.cost_pack, nullptr, workspace, stream);
  if (rc) return rc;
  plan_begin_kernel<<<1, 32, 0, st>>>(s.st, s.cost_pack);
  for (int idx = 0; idx < num_iters; ++idx) {
    g_eval_gate = &s.st->active;
    rc = klerg_eval_gradient(k, dyn, bar, peers, x0, R0, u, H, packed, N, ld, q_base, p, p_stats, floor, Rinv_diag, alpha,
                             ctrl_lo, ctrl_hi, s.v_grad, nullptr, s.totals, nullptr, s.dgdx, s.du, s.djdlam, s.u_star, nullptr,
                             nullptr, workspace, stream);
    g_eval_gate = nullptr;
    if (rc) return rc;
    plan_select_kernel<<<1, 256, 0, st>>>(s.st, idx, (int)H, A, s.djdlam, s.u_star, u, s.cands, fixed_lam, lam, max_app_dur);
    g_eval_gate = &s.st->n_cand;
    rc = klerg_eval_costs(k, dyn, bar, peers, x0, R0, s.cands, ngate, H, packed, N, ld, q_base, p, p_stats, floor, s.v_costs,
                          nullptr, nullptr, s.cost_pack, nullptr, workspace, stream);
    g_eval_gate = nullptr;
    if (rc) return rc;
    plan_accept_kernel<<<1, 256, 0, st>>>(s.st, idx, (int)H, A, s.cost_pack, s.cands, u, fixed_lam);
  }
  plan_finish_kernel<<<1, 256, 0, st>>>(s.st, (int)(H * A), u, result, ws_fused_ctrl(workspace));
  if (int e = check_launch("plan kernels")) return e;
  return klerg_rollout(dyn, nullptr, x0, R0, u, 1, H, result + 8 + H * A, nullptr, nullptr, nullptr, nullptr, stream);
}

/*
 * klerg_b200.h -- C ABI of the B200-native KL-ergodic hot path (libklerg_b200.so).
 *
 * Drop-in boundary for franka_test/scripts/control_torch of
 * apinosky/embodied-active-learning-vision.  The reference has no native
 * interface for this path (it is torch-CPU Python), so each
 * entry point cites the reference *function* whose arithmetic it replaces
 * (paths relative to franka_test/scripts/control_torch).  The Python mirror in
 * embodied-active-learning-vision_b200/control_torch binds these with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the parameter is a `const klerg_*_spec*` (small host struct, read at
 *     call time) or is documented as host memory;
 *   - nothing is allocated inside: outputs and `workspace` are caller-owned;
 *     `workspace` must be at least klerg_workspace_bytes() bytes, zero-filled
 *     once at allocation (kernels leave it zeroed where they need zeros);
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as
 *     void*); no call synchronises the device;
 *   - return value 0 = success, negative = error (see klerg_last_error());
 *     no exceptions cross the boundary;
 *   - fp32 data, row-major; indices int64 (torch.randperm's dtype);
 *   - reentrant per stream, not thread-safe per workspace.
 */
#ifndef KLERG_B200_H
#define KLERG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KLERG_MAX_D 8   /* explored dimensions        (reference uses 2..6) */
#define KLERG_MAX_S 24  /* planner state dimension    (12 for xyzrpw)       */
#define KLERG_MAX_A 8   /* controls                                         */
#define KLERG_MAX_H 256 /* planning horizon                                 */

/* Gaussian-like pairwise kernel psi (klerg_utils.py:7-10). `scale[d]` is the
 * reference's `std[d]`: it divides the squared difference UN-squared and its
 * absolute value is used (klerg_utils.py:20,27,13). */
typedef struct klerg_kernel_spec {
  int32_t D;                  /* len(explr_idx)                              */
  int32_t S;                  /* columns per state row                        */
  int32_t explr[KLERG_MAX_D]; /* explored columns of a state row              */
  float scale[KLERG_MAX_D];
  float nu;                   /* psi is divided by nu                         */
} klerg_kernel_spec;

enum { KLERG_DYN_SINGLE = 0, KLERG_DYN_DOUBLE = 1, KLERG_DYN_SPEED = 2, KLERG_DYN_ROLL = 3 };

/* Integrator models of dynamics.py (SingleIntegratorEnv :67, DoubleIntegratorEnv
 * :81, DoubleIntegratorSpeedEnv :97, DoubleIntegratorRollEnv :224). */
typedef struct klerg_dyn_spec {
  int32_t kind;
  int32_t S; /* num_states  */
  int32_t A; /* num_actions */
  float dt;
  int32_t rpw[3];      /* ROLL: indices of roll,pitch,yaw among the positions  */
  int32_t has_ang_map; /* ROLL: affine rot<->angle map (klerg.py:147-149)      */
  float rot_lo[3], rot_hi[3]; /* "robot_lim" side of the map                   */
  float ang_lo[3], ang_hi[3]; /* "tray_lim" (real angle) side of the map       */
} klerg_dyn_spec;

/* BarrierFunction (barrier.py:40-90): limits already shrunk by b_buff. n = 0
 * stands for NoBarrier (barrier.py:147-159). */
typedef struct klerg_barrier_spec {
  int32_t n;
  float lo[KLERG_MAX_S], hi[KLERG_MAX_S];
  float weight[KLERG_MAX_S], power[KLERG_MAX_S];
} klerg_barrier_spec;

const char* klerg_last_error(void);
int klerg_abi_version(void);
/* Number of kernel launches issued through this library since load (bench accounting). */
long long klerg_launch_count(void);
/* Peak-rate probe for the roofline denominators: `blocks` CTAs x 256 threads each
 * run iters*64 dependent-chain-of-8 FFMA (kind 0), MUFU.EX2 (kind 1) or packed
 * FFMA2 (kind 2: two fp32 lanes per instruction) ops. */
int klerg_peak_probe(int kind, int iters, int blocks, float* out, void* stream);
/* SM count / compute capability of the current device (host out-pointers). */
int klerg_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Upper bound of workspace bytes for G segments (candidates/targets). */
size_t klerg_workspace_bytes(int64_t G);

/* ---- workspace samples --------------------------------------------------- */

/* AoS samples [N,D] (layout of env_sampler.sample, klerg.py:375) -> packed SoA
 * [D][ld] pre-multiplied by sqrt(0.5*log2(e)/|scale_d|) so that
 * psi = 2^-(sum_d (t'_d - s'_d)^2).  ld >= N, ld % 4 == 0. */
int klerg_pack_samples(const klerg_kernel_spec* k, const float* samples, int64_t N, float* packed,
                       int64_t ld, void* stream);

/* ---- a2 / a3: traj_footprint_vec, traj_spread_vec (klerg_utils.py:17-29) -- */

/* For every segment g < G and sample i < N:
 *   out[g*out_stride + i] = (add_in ? add_in[i] : 0) + red_j psi(states[g][j], s_i)
 * red = sum (mode 0) or max (mode 1), j < T, states[g] = states + g*seg_stride
 * (rows of k->S floats).  totals[g*2+{0,1}] = {sum_i out, max_i out} (this
 * rank's samples only).  T = 0 yields add_in (or zeros), as klerg.py:497-498. */
int klerg_footprint(const klerg_kernel_spec* k, int mode, const float* states, int64_t G, int64_t T,
                    int64_t seg_stride, const float* packed, int64_t N, int64_t ld, const float* add_in,
                    float* out, int64_t out_stride, double* totals, void* workspace, void* stream);

/* Both reductions in ONE pass over the squared distances, for the two memory-buffer passes of a planner step that
 * visit the same rows (klerg.py:470-475 spread of get_target_dist over ALL buffer rows; klerg.py:496 footprint of
 * the drawn rows): states[T][S] hold the drawn rows FIRST (T_sum of them, any order) and the remaining buffer rows
 * behind them;  out_sum[i] = sum_{j < T_sum} psi(states[j], s_i),  out_max[i] = max_{j < T} psi(states[j], s_i)
 * (both [ld]); totals[0..1] = {sum_i, max_i} of out_sum over this rank's samples. */
int klerg_footprint_sum_max(const klerg_kernel_spec* k, const float* states, int64_t T, int64_t T_sum,
                            const float* packed, int64_t N, int64_t ld, float* out_sum, float* out_max,
                            double* totals, void* workspace, void* stream);

/* The same pass with the squared distances on the tensor cores: e_ij = |sc_i|^2 + |xc_j|^2 - 2 xc_j . sc_i is a
 * bilinear form of D + 2 <= 8 terms = one K = 8 step of tcgen05.mma kind::tf32 (3xTF32 split, fp32 accumulation,
 * samples as TMEM lanes, 128 state rows per accumulator); the CUDA cores keep min / exp / add, so the pass is bound
 * by the MUFU pipe (one exp per pair) instead of the FP32 pipe.  Around the centre of the states' bounding box; when a
 * state lies outside the radius of the expanded pair form (klerg_pair.cuh) the launch falls back, on the device, to
 * the CUDA-core pass of klerg_footprint_sum_max.  D <= 6, T >= 1, N >= 1.  scratch: klerg_footprint_tc_scratch_bytes(T)
 * bytes of device memory, 128-byte aligned (packed state chunks). */
int64_t klerg_footprint_tc_scratch_bytes(int64_t T);
int klerg_footprint_sum_max_tc(const klerg_kernel_spec* k, const float* states, int64_t T, int64_t T_sum,
                               const float* packed, int64_t N, int64_t ld, float* out_sum, float* out_max,
                               double* totals, void* workspace, void* scratch, int64_t scratch_bytes, void* stream);

/* a1: psi_fn / dpsi_dx_fn (klerg_utils.py:7-15) materialised: psi[N][T] = psi(states_j[explr], samples_i)
 * and / or dpsi[N][T][D] = -(x_j - s_i)/|scale| * psi * nu (dpsi_dx_fn divides by nu once, through psi).
 * `samples` are the raw AoS samples [N][D].  Either output may be NULL. */
int klerg_psi_matrix(const klerg_kernel_spec* k, const float* states, int64_t T, const float* samples, int64_t N,
                     float* psi, float* dpsi, void* stream);

/* ---- a5 / a6: renormalize, cost_norm (klerg_utils.py:38-58) --------------- */

/* stats[0..3] = {sum, max, min, #nan} of x[0..N). */
int klerg_vector_stats(const float* x, int64_t N, double* stats, void* workspace, void* stream);
/* out = c / max(c), c = max(x/sum, floor) -- closed form of renormalize(). */
int klerg_renormalize(const float* x, int64_t N, float floor, float* out, void* workspace, void* stream);
/* Same with externally supplied (e.g. rank-combined) stats = {sum, max} on the device. */
int klerg_renormalize_with_stats(const float* x, int64_t N, const double* stats, float floor, float* out,
                                 void* stream);
/* in place: NaN -> 1e-6, then x /= sum(x). */
int klerg_cost_norm(float* x, int64_t N, void* workspace, void* stream);

/* ---- a4: kldiv_grad_vec (klerg_utils.py:12-15, 31-36) -------------------- */

/* dgdx[t][explr[d]] = sum_i w_i * (-(x_td - s_id)/|scale_d|) * psi(x_t, s_i)
 * for t < H (the reference loops t on the host, klerg.py:440-443); other
 * columns of dgdx[H][S] are zero.  Explicit importance ratio w[N]. */
int klerg_kl_gradient(const klerg_kernel_spec* k, const float* states, int64_t H, const float* packed,
                      int64_t N, int64_t ld, const float* w, float* dgdx, void* workspace, void* stream);

/* Planner form: the importance ratio p/q of klerg.py:436 is formed on the fly
 * from v = q_base + q_iter (output of klerg_footprint), its global totals
 * (`world` rank blocks of [sum,max], summed/maxed in rank order) and p:
 *   c_i = max(v_i/sum v, floor), q_i = c_i/max c, w_i = p_i/q_i.
 * Writes this rank's partial gradient grad_part[H][D] (explored dims only,
 * doubles) and kl_part[2] = {sum_i p_i (log p_i - log c_i), sum_i c_i}. */
int klerg_kl_gradient_fused(const klerg_kernel_spec* k, const float* states, int64_t H,
                            const float* packed, int64_t N, int64_t ld, const float* v,
                            const double* totals, int world, const float* p, float floor,
                            double* grad_part, double* kl_part, void* workspace, void* stream);

/* ---- a11: KL(p||q) cost of get_cost (klerg.py:686-710) -------------------- */

/* Per candidate g: kl_part[g*2+{0,1}] = {sum_i p_i (log p_i - log c_i), sum_i c_i}
 * over this rank's samples, c from v[g] and totals[world][G][2]. */
int klerg_kl_cost_partial(const float* v, int64_t v_stride, int64_t G, int64_t N, const double* totals,
                          int world, const float* p, float floor, double* kl_part, void* workspace,
                          void* stream);
/* cost[g] = Sa/sum_p - log(sum_p) + log(Sc) + barrier_sum[g], with {Sa,Sc}
 * summed over `world` blocks of kl_part[G][2] and sum_p = p_stats[0]. */
int klerg_kl_cost_final(const double* kl_part, int world, int64_t G, const double* p_stats,
                        const float* barrier_sum, float* cost, void* stream);

/* ---- a7: get_target_dist weighting (klerg.py:452-486) --------------------- */

/* stage 1: acc[0..3] = {max_i spread_i, sum_{i inside} spread_i, #outside, min_i p_i}
 * `samples` are the raw AoS samples, lim_lo/hi (host, D floats) = robot_lim
 * rows of the explored dims (klerg.py:454).  spread may be NULL (empty buffer:
 * klerg.py:476-477) in which case acc = {1, 0, 0, min p}. */
int klerg_target_stage1(const float* samples, int32_t D, int64_t N, const float* lim_lo,
                        const float* lim_hi, const float* spread, const float* p, double* acc,
                        void* workspace, void* stream);
/* Rank combine for sharded runs: blocks = `world` rank blocks of n doubles,
 * kinds[q] (host) = 0 sum, 1 max, 2 min; out[q] = reduction over ranks. */
int klerg_combine_blocks(const double* blocks, int world, int n, const int* kinds, double* out, void* stream);
/* expo = {mean_i spread'_i (klerg.py:474-481), min p, max spread} from the
 * (rank-combined) stage-1 acc and the global sample count. */
int klerg_target_exponent(const double* acc, int64_t N_total, double* expo, void* stream);
/* stage 2: mode 0 (weight_temp / plot): p2 = p ** expo[0];
 *          mode 1 (weight_env): p2 = p + (1 - spread'_i) * expo[1], spread' as klerg.py:474-475;
 *          mode 2: p2 = p.   acc[0..1] = {sum p2, max p2}.  expo = device doubles
 *          {exponent, min p, max spread}. */
int klerg_target_stage2(int mode, const float* samples, int32_t D, int64_t N, const float* lim_lo,
                        const float* lim_hi, const float* spread, const float* p, const double* expo,
                        float* p2, double* acc, void* workspace, void* stream);
/* stage 3: p = (renormalize(p2)) ** temp given acc2 = {sum p2, max p2} (device);
 * `renorm` = 0 skips the renormalize (no weighting flags set).  p_stats[0] = sum p. */
int klerg_target_stage3(const float* p2, int64_t N, const double* acc2, int renorm, float floor, float temp,
                        float* p, double* p_stats, void* workspace, void* stream);

/* ---- a9/a11/a15/a16-a18: rollout, barrier, linearisation ------------------ */

/* One thread per candidate b < B rolls u[b][H][A] out from x0[S] (and R0[9] for
 * ROLL, row-major) with RK4 (dynamics.py:7-13,58-65):
 *   traj[b][0] = x0, traj[b][t+1] = step(traj[b][t], u[b][t])      [B][H+1][S]
 *   barrier_sum[b] = sum_t barr(traj[b][t+1])                     (klerg.py:708)
 *   dbarr[b][t]    = dbarr(traj[b][t])            [B][H][S]       (klerg.py:425), may be NULL
 *   P[b][t]        = d(pos rate)/d(vel) block of A_t  [B][H][A*A]  (klerg.py:423), may be NULL
 *   R_out[b]       = rotation matrix after the last step [B][9], may be NULL. */
int klerg_rollout(const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar, const float* x0,
                  const float* R0, const float* u, int64_t B, int64_t H, float* traj,
                  float* barrier_sum, float* dbarr, float* P, float* R_out, void* stream);

/* ---- a20/a22: state-feedback default policies (default_policies.py:53-119) ----
 * Robot.forward (klerg.py:409-431) evaluates u_t = policy(x_t) while it rolls the planner out; with the Roll / Zero
 * policies that is the stored plan and dmu/dx = 0 (every other entry point of this header).  BarrierPush and LQR
 * feed the state back:
 *   KLERG_POLICY_LQR           u_t = -K x_t, dmu/dx = -K                              (default_policies.py:100-119)
 *   KLERG_POLICY_BARRIER_PUSH  u_t = u_in[t] (use_u != 0) or 0, then for every position i with
 *                              (x_i >= 1 and v_i > 0) or (x_i <= -1 and v_i < 0): u_i = -weight v_i,
 *                              dmu_i/dv_i = -weight                                   (default_policies.py:53-97)
 * klerg_policy_rollout runs the closed loop (serial in t: one warp) and returns the controls it applied,
 * u_eff [H][A], and dmudx [H][A][S]; the open-loop entry points then take u_eff (same trajectory), and
 * klerg_adjoint_policy is the adjoint sweep of klerg.py:433-450 with the closed-loop linearisation
 * rho' = dgdx_t - dbarr_t - (A_t + B dmudx_t)^T rho, one RK4 step of -dt per t (SINGLE, DOUBLE and ROLL models). */
enum { KLERG_POLICY_LQR = 1, KLERG_POLICY_BARRIER_PUSH = 2 };
typedef struct klerg_policy_spec {
  int32_t kind;
  int32_t use_u;                        /* BARRIER_PUSH: replay u_in (iter_idx > 0) or start from zeros        */
  float weight;                         /* BARRIER_PUSH: 5 in the reference                                   */
  float K[KLERG_MAX_A * KLERG_MAX_S];   /* LQR gain, row-major [A][S]                                         */
} klerg_policy_spec;
int klerg_policy_rollout(const klerg_dyn_spec* dyn, const klerg_policy_spec* pol, const float* x0, const float* R0,
                         const float* u_in, int64_t H, float* u_eff, float* dmudx, void* stream);
/* dgdx, dbarr [H][S]; P [H][A*A] or NULL (0.8 I); dmudx [H][A][S]; u [H][A] = u_eff.  Host arrays as klerg_adjoint. */
int klerg_adjoint_policy(const klerg_dyn_spec* dyn, int64_t H, const float* dgdx, const float* dbarr, const float* P,
                         const float* dmudx, const float* u, const float* Rinv_diag, float alpha,
                         const float* ctrl_lo, const float* ctrl_hi, float* du, float* djdlam, float* u_star,
                         void* stream);

/* barr(x_t) and dbarr(x_t) for T rows of S floats (barrier.py:70-87). */
int klerg_barrier_eval(const klerg_barrier_spec* bar, const float* x, int64_t T, int32_t S,
                       float* value, float* grad, void* stream);

/* The two barrier variants the reference keeps beside BarrierFunction (never instantiated by its Robot):
 *   x_ref != NULL  VelocityBarrier (barrier.py:162-205): the limits of row t are x_ref[t] + {lo, hi} (a band around
 *                  the other state; weight 0 on the rows the band does not apply to);
 *   tilt != NULL   TiltBarrierFunction (barrier.py:95-144): tilt = acos(cos r cos p) of the (mapped) roll / pitch
 *                  angles; the yaw limits become tilt/pi * {w_lo, w_hi}; value += [tilt <= tilt_lim] weight
 *                  (tilt - tilt_lim)^power, and its derivative lands on the r and p columns (no chain factor for the
 *                  map, as in the reference).  tilt_out [T] (optional) receives the tilt of every row (the reference
 *                  leaves the yaw limits of the LAST evaluated row in the wrapped barrier). */
typedef struct klerg_tilt_spec {
  int32_t r_idx, p_idx, w_idx, has_map;
  float w_lo, w_hi;
  float tilt_lim, power, weight;
  float rot_lo[2], rot_hi[2], ang_lo[2], ang_hi[2]; /* affine map of r, p from state units to angles */
} klerg_tilt_spec;
int klerg_barrier_eval_ext(const klerg_barrier_spec* bar, const float* x, const float* x_ref, const klerg_tilt_spec* tilt,
                           int64_t T, int32_t S, float* value, float* grad, float* tilt_out, void* stream);

/* ---- a10: adjoint sweep of Robot.backward (klerg.py:433-450, 590-593) ----- */

/* grad_part: `world` blocks of [H][D] doubles (klerg_kl_gradient_fused), summed
 * in rank order.  rho_H = 0; for t = H-1..0 one RK4 step of
 * rho' = dgdx_t - dbarr_t - A_t^T rho with step -dt (default policy: dmudx = 0),
 * du_t = -Rinv B_t^T rho, djdlam_t = rho B_t du_t, u_star = clamp(u + alpha du).
 * P may be NULL (0.8*I).  Rinv_diag, ctrl_lo, ctrl_hi are HOST arrays of A
 * floats (diag of R^-1, control_lim columns, klerg.py:193,197).
 * Outputs: dgdx[H][S], du[H][A], djdlam[H], u_star[H][A]. */
int klerg_adjoint(const klerg_dyn_spec* dyn, const klerg_kernel_spec* k, int64_t H,
                  const double* grad_part, int world, const float* dbarr, const float* P,
                  const float* traj, const float* u, const float* Rinv_diag, float alpha,
                  const float* ctrl_lo, const float* ctrl_hi, float* dgdx, float* du, float* djdlam,
                  float* u_star, void* stream);

/* ---- fused evals: one launch per eval -------------------------------------- */

/* Process-wide switches of the fused evals.
 *   KLERG_OPT_EVAL_OVERLAP  1 = consecutive fused evals on a stream are independent (no eval reads what the
 *                           previous one wrote): the next eval's CTAs do not wait for the previous eval's last
 *                           CTA (gather + adjoint) to finish.  Default 0: stream order as usual.
 *   KLERG_OPT_GRID_LIMIT    > 0: at most this many CTAs per fused launch (tests).
 *   KLERG_OPT_PDL           0 = plain cooperative launches without programmatic dependent launch.
 *   KLERG_OPT_COOP_WITH_PDL (read-only) 1 / 0 / -1: the driver accepts both attributes / not / not probed yet.
 *   KLERG_OPT_EXACT_PAIRS   1 = every pair pass (footprint, fused evals) evaluates |x - s|^2 from the coordinate
 *                           differences; default 0: the cheaper expanded form |sc|^2 + |xc|^2 - 2 xc.sc around a centre
 *                           of the state set wherever the set is narrow enough for it (relative error of psi < 3e-5),
 *                           the difference form elsewhere.
 *   KLERG_OPT_MIXED_WARPS   schedule of the gradient pass for D >= 5 (A/B switch).  0 (default): 12 warps with 4 states
 *                           each and the <= 2 states that remain shared over all warps, where the horizon allows
 *                           (H = 48..50), else 16 warps with 3-4 states; 1: always the 16-warp split; 16: 16 warps
 *                           with 3 states each + shared rest (spills at 128 registers: slower; kept for measurement).
 *   KLERG_OPT_SATURATE_MILLI  0 (default): u* = clamp(u + alpha du, ctrl_lo, ctrl_hi) (klerg.py:522); t > 0: the
 *                           reference's `saturate` flag, u* = tanh((u + alpha du) / (t / 1000)) * ctrl_hi
 *                           (Robot.saturate_control, klerg.py:342-349; the reference uses t = 100).  Read when an
 *                           eval / adjoint launch is enqueued. */
enum { KLERG_OPT_EVAL_OVERLAP = 1, KLERG_OPT_GRID_LIMIT = 2, KLERG_OPT_PDL = 3, KLERG_OPT_COOP_WITH_PDL = 4,
       KLERG_OPT_EXACT_PAIRS = 5, KLERG_OPT_MIXED_WARPS = 6, KLERG_OPT_SATURATE_MILLI = 7 };
int klerg_set_option(int key, int value);
int klerg_get_option(int key);
/* Two ranks on ONE GPU (tests of the exchange protocol on a single-GPU box): after klerg_emu_begin() the next
 * klerg_eval_gradient* / klerg_eval_costs call of rank 0 and of rank 1 (peers->world == 2, both mailboxes on this
 * device) are recorded instead of launched; klerg_emu_launch() runs them as one cooperative grid whose halves
 * act as the two ranks. */
int klerg_emu_begin(void);
int klerg_emu_launch(void* stream);

/* Sample-sharding peers (samples of one workspace split over the GPUs of a
 * box).  mailbox[r] is rank r's exchange buffer
 * (klerg_mailbox_bytes() bytes, zero-filled once) mapped into THIS process -
 * peer-to-peer device memory reachable over NVLink; mailbox[rank] is the local
 * one.  NULL / world <= 1 = single GPU.  All ranks must issue the same
 * sequence of fused evals on a mailbox set. */
typedef struct klerg_peers {
  int32_t world, rank;
  void* mailbox[8];
  int64_t n_max; /* samples of the largest shard: every rank sizes its grid from it, so that all ranks
                  * launch the same number of CTAs (0 = this rank's N) */
} klerg_peers;
size_t klerg_mailbox_bytes(void);
/* Mailbox plumbing for one-process-per-GPU ranks (CUDA IPC): create allocates and zero-fills this
 * rank's mailbox and returns its 64-byte IPC handle (host memory) to be sent to the peers by any
 * means (e.g. torch.distributed.all_gather_object); open maps a peer's mailbox; close unmaps /
 * frees.  The reference has no equivalent (single process, CPU). */
int klerg_mailbox_create(void** ptr, unsigned char* handle64);
int klerg_mailbox_open(const unsigned char* handle64, void** ptr);
int klerg_mailbox_close(void* ptr, int owner);
/* Byte offset inside `workspace` of 8 int64 SM-cycle stamps left by the last klerg_eval_gradient
 * (start, rollout, forward, meet-1, gradient, meet-2, reduce, end; relative to start). */
size_t klerg_debug_stamps_offset(void);
/* KLERG_STAMPS builds, single GPU: byte offset inside `workspace` of [160 CTAs][8] uint64 globaltimer stamps (ns) of the
 * last klerg_eval_gradient: start, rollout, forward, meeting 1, gradient, entry sums, (finisher) end. */
size_t klerg_debug_cta_stamps_offset(void);
/* Byte offset inside `workspace` of a sticky uint32 fault word: non-zero after a fused eval gave
 * up waiting at a meeting point (results of that eval are undefined). */
size_t klerg_fused_fault_offset(void);

/* One planner iteration up to the control update (klerg.py:505-523):
 *   forward(): rollout of u[H][A] from x0 (pre-step states, linearisation, dbarr)
 *   q = renormalize(q_base + footprint(traj))          (sum/max over ALL ranks)
 *   backward(): dgdx_t = kldiv_grad_vec(x_t, samples, p/q), adjoint sweep
 * -> dgdx[H][S], du[H][A], djdlam[H], u_star = clamp(u + alpha du)[H][A].
 * Also: traj[H+1][S] (may be NULL), totals[2] = {sum, max} of q_base + q_iter,
 * kl_out[2] = {sum_i p_i (log p_i - log c_i), sum_i c_i}, cost[1] = KL of that
 * footprint + barrier sum (each may be NULL).  p must hold ld floats (entries >= N are ignored; whole tile
 * rows are copied with TMA) and be 16-byte aligned.  v_scratch: ld floats (receives
 * q_base + q_iter of this rank's samples).  Rinv_diag / ctrl_lo / ctrl_hi are
 * HOST arrays of A floats.  Replaces klerg_rollout + klerg_footprint +
 * klerg_kl_gradient_fused + klerg_adjoint with a single launch.
 * fault_out (may be NULL): one float next to the outputs the host reads anyway, 1.0 if an in-kernel wait of
 * this workspace ever timed out (copy of the sticky fault word), else 0.0. */
int klerg_eval_gradient(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                        const klerg_peers* peers, const float* x0, const float* R0, const float* u, int64_t H,
                        const float* packed, int64_t N, int64_t ld, const float* q_base, const float* p,
                        const double* p_stats, float floor, const float* Rinv_diag, float alpha,
                        const float* ctrl_lo, const float* ctrl_hi, float* v_scratch, float* traj,
                        double* totals, float* cost, float* dgdx, float* du, float* djdlam, float* u_star,
                        double* kl_out, float* fault_out, void* workspace, void* stream);

/* The same for K <= 32 belief targets p_k over one workspace and one trajectory (fingerprint test mode,
 * test_fingerprint_main.py swaps robot.target_dist between beliefs; BASELINE config 5): the rollout, the
 * forward pair pass and q are shared, the gradient pair pass / reduction / adjoint run once per target inside
 * the same launch.  p[K][p_stride] (p_stride >= N, multiple of 4), p_stats[K]; outputs are [K][...] stacked:
 * cost[K], dgdx[K][H][S], du[K][H][A], djdlam[K][H], u_star[K][H][A], kl_out[K][2]. */
int klerg_eval_gradient_targets(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                                const klerg_peers* peers, const float* x0, const float* R0, const float* u, int64_t H,
                                const float* packed, int64_t N, int64_t ld, const float* q_base, const float* p,
                                int64_t K, int64_t p_stride, const double* p_stats, float floor,
                                const float* Rinv_diag, float alpha, const float* ctrl_lo, const float* ctrl_hi,
                                float* v_scratch, float* traj, double* totals, float* cost, float* dgdx, float* du,
                                float* djdlam, float* u_star, double* kl_out, float* fault_out, void* workspace,
                                void* stream);

/* Robot.get_cost (klerg.py:686-710) for G <= 8 candidate control sequences
 * u[G][H][A] in one launch (the line-search windows of klerg.py:712-751):
 * cost[g] = KL(p || renormalize(q_base + footprint(post-step states))) + barrier.
 * v_scratch: G*ld floats; traj[G][H+1][S] and totals[G][2] may be NULL. */
int klerg_eval_costs(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                     const klerg_peers* peers, const float* x0, const float* R0, const float* u, int64_t G,
                     int64_t H, const float* packed, int64_t N, int64_t ld, const float* q_base, const float* p,
                     const double* p_stats, float floor, float* v_scratch, float* traj, double* totals,
                     float* cost, float* fault_out, void* workspace, void* stream);

/* The optimisation loop of Robot.kldiv_planner (klerg.py:505-576) with its line search (klerg.py:712-751), enqueued
 * at once and decided on the device:  cost(u) | { gradient(u) -> argmin djdlam, line-search windows -> costs of the
 * candidates -> accept rules } x num_iters | nan_to_num(u), rollout(u).  The evals are the launches above, gated by
 * a device-side loop state: after the reference's `break` the remaining ones return immediately, and a cost launch
 * evaluates exactly the candidates the host loop would pass.  Same float32 comparisons in the same order as the host
 * code.  u[H][A] (device) is the current plan on entry and the optimised plan on return.  fixed_lam != 0: the
 * window [t_app, t_app + lam) instead of the line search.  max_app_dur: 5 in the reference (klerg.py:712).
 * result (device, klerg_plan_result_floats(H, S, A) floats, the ONE buffer the host reads back):
 *   {last_cost, fault, cost evals, gradient evals, accepted iterations, 0, 0, 0, u[H][A], traj[H+1][S] of rollout(u)}.
 * scratch: klerg_plan_scratch_bytes(H, S, A, ld) bytes (device). */
size_t klerg_plan_scratch_bytes(int64_t H, int32_t S, int32_t A, int64_t ld);
int64_t klerg_plan_result_floats(int64_t H, int32_t S, int32_t A);
int klerg_plan_optimize(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                        const klerg_peers* peers, const float* x0, const float* R0, float* u, int64_t H,
                        const float* packed, int64_t N, int64_t ld, const float* q_base, const float* p,
                        const double* p_stats, float floor, const float* Rinv_diag, float alpha, const float* ctrl_lo,
                        const float* ctrl_hi, int32_t num_iters, int32_t fixed_lam, int32_t lam, int32_t max_app_dur,
                        void* scratch, float* result, void* workspace, void* stream);

/* The same for ANY number of candidates B (BASELINE config 3: 1024 candidates x horizon 50 x 1e6 samples) in ONE launch:
 * rollouts spread over the grid, forward pair pass per group of 8 candidates over L2-resident samples, one grid-wide
 * meeting for all candidates' normalisers, KL pass, costs.  Single GPU.  v_scratch: B*ld floats; scratch:
 * klerg_eval_costs_batch_scratch_bytes(B, H, D) bytes; totals[B][2] may be NULL. */
size_t klerg_eval_costs_batch_scratch_bytes(int64_t B, int64_t H, int32_t D);
int klerg_eval_costs_batch(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                           const float* x0, const float* R0, const float* u, int64_t B, int64_t H, const float* packed,
                           int64_t N, int64_t ld, const float* q_base, const float* p, const double* p_stats,
                           float floor, float* v_scratch, void* scratch, double* totals, float* cost, float* fault_out,
                           void* workspace, void* stream);

/* ---- a8: workspace samples of Robot.get_samples (klerg.py:173,375) --------- */

/* samples[i][d] = low[d] + torch.rand(n_rows, D)[i][d] * high_minus_low[d] for rows row_lo <= i < row_hi, drawn with
 * torch's CPU generator CONTINUED ON THE DEVICE: state624 (device, 624 words) / left / next are the MT19937 state
 * and counters of torch.get_rng_state() (ATen CPUGeneratorImplStateLegacy: left int32 at byte 8, next uint64 at 16,
 * state uint64[624] at 24).  The whole stream of n_rows * D numbers is generated (bit-exact with the host draw);
 * only the requested rows are written: out[(i - row_lo) * D + d].  state_out (device, 626 words) receives the state
 * followed by left and next after the draw - the caller writes them back with torch.set_rng_state so that the
 * memory-buffer randperm that follows (memory_buffer.py:58) draws what it would have drawn.  low / high_minus_low:
 * HOST arrays of D floats (Uniform.low, high - low as torch computes it in fp32). */
int klerg_mt19937_uniform(const uint32_t* state624, int32_t left, int32_t next, int64_t n_rows, int32_t D,
                          const float* low, const float* high_minus_low, int64_t row_lo, int64_t row_hi, float* out,
                          uint32_t* state_out, void* stream);

/* ---- a19: memory-buffer selection (memory_buffer.py:52-63) ---------------- */

/* out[m] = table[idx[m]] for m < M; idx are the host-drawn torch.randperm
 * indices (int64).  Returns -3 via a device flag check only in debug builds. */
int klerg_gather_rows(const float* table, int32_t S, const int64_t* idx, int64_t M, float* out,
                      void* stream);

/* klerg_adjoint for K belief targets sharing one trajectory (one CTA per target): grad_parts[K][world][H][D],
 * outputs dgdx[K][H][S], du[K][H][A], djdlam[K][H], u_star[K][H][A]. */
int klerg_adjoint_targets(const klerg_dyn_spec* dyn, const klerg_kernel_spec* k, int64_t H, int64_t K,
                          const double* grad_parts, int world, const float* dbarr, const float* P,
                          const float* traj, const float* u, const float* Rinv_diag, float alpha,
                          const float* ctrl_lo, const float* ctrl_hi, float* dgdx, float* du, float* djdlam,
                          float* u_star, void* stream);

/* ---- a4 for K belief targets: shared-psi contraction on the tensor cores ---- */

/* kldiv_grad_vec (klerg_utils.py:12-15,31-36) for all H states and K targets p_k over one workspace and one
 * trajectory (BASELINE config 5): same inputs and outputs as klerg_kl_gradient_fused, per target -
 *   grad_part[k][H][D] (doubles, this rank's partial), kl_part[k][2] = {sum_i p_ki (log p_ki - log c_i), sum_i c_i}
 * - but psi(x_t, s_i) is evaluated once per state-sample pair and the sum over the samples runs as a tf32 tensor-core
 * contraction (3xTF32 split, fp32 accumulate).  kl_part may be NULL (skips a log per target and sample).
 * Limits: H <= 64, K <= 32, K*(D+1) <= 128.  `scratch`:
 * klerg_kl_gradient_targets_scratch_bytes(H, K) bytes; `fault` (may be NULL) is set to 1 if an in-kernel wait timed out. */
size_t klerg_kl_gradient_targets_scratch_bytes(int64_t H, int64_t K);
int klerg_kl_gradient_targets(const klerg_kernel_spec* k, const float* states, int64_t H, const float* packed,
                              int64_t N, int64_t ld, const float* v, const double* totals, int world,
                              const float* P, int64_t K, int64_t p_stride, float floor, double* grad_part,
                              double* kl_part, void* scratch, uint32_t* fault, void* stream);

/* ---- fingerprint belief update ----------------------------- */

/* FingerprintDist.update_prior (franka_test/scripts/dist_modules/fingerprint_module.py:539-589; meas_footprint_vec
 * :417-424; numpy renormalize, control/klerg_utils.py:41-54): the per-grid-point normal belief {prior, prior_var}
 * over grid[G][D] (the 50^D mesh of build_grid, :504-519) after n measurements at locs[n][D].  meas_sum =
 * sum_j (val_j / 2 + 0.5) over the processed measurement values (process_meas, :470-478; a host scalar).  All arrays
 * DEVICE float64 like the reference's numpy arrays; posterior / posterior_var may alias prior / prior_var.
 * scratch: klerg_belief_scratch_bytes(G, n) bytes. */
size_t klerg_belief_scratch_bytes(int64_t G, int32_t n);
int klerg_belief_update(const double* grid, int64_t G, int32_t D, const double* locs, int32_t n, double scale,
                        double meas_sum, const double* prior, const double* prior_var, double* posterior,
                        double* posterior_var, void* scratch, void* stream);

/* ---- target density of the VAE sensor model -------------- */

/* VAE.pdf_torch (franka_test/scripts/vae/vae.py:244-275), the `target_dist`
 * duck type Robot.get_target_dist calls (klerg.py:463):
 *   latent = [z || s - shift];  y = decode(latent)[:, :n_logvar]
 *   p(s) = amax_l exp( mean_z clamp(y_l, clamp_lo, clamp_hi) )
 * for decode = Linear(z_dim+s_dim, h1) ReLU Linear(h1, h2) ReLU Linear(h2, out)
 * (vae.py:77-84; hidden_dim [512,256] -> h1 = 256, h2 = 512).  n_z = 1 is the
 * default `z_samples` row (vae.py:258-259); n_z > 1 the z buffer (vae.py:253-257,
 * 268-270).  Limits: s_dim 1..7, h1 % 8 == 0 (<= 1024), h2 % 32 == 0 up to 256 or
 * h2 % 64 == 0 up to 512, n_logvar 1..15.
 *
 * klerg_target_decoder_pack folds z into the first-layer bias and splits W2 into
 * tf32 hi/lo operand stages; call it whenever the weights or z change (all
 * pointers DEVICE, torch.nn.Linear layouts: w1[h1][z_dim+s_dim], w2[h2][h1],
 * w3[>=n_logvar][h2], z[n_z][z_dim]).  `packed`: klerg_target_decoder_packed_bytes()
 * bytes, 128-byte aligned.
 *
 * klerg_target_decoder_pdf evaluates p for samples[N][s_dim] (raw AoS samples,
 * the argument of pdf_torch); `shift` = seed_x when the model was built with
 * dx=True (vae.py:249-250), else NULL.  The second layer runs on the tensor
 * cores (tcgen05.mma kind::tf32, 3xTF32 split, fp32 accumulate in TMEM).
 * `fault` (may be NULL): set to 1 if an in-kernel wait timed out. */
size_t klerg_target_decoder_packed_bytes(int32_t s_dim, int32_t z_dim, int32_t n_z, int32_t h1, int32_t h2,
                                         int32_t n_logvar);
int klerg_target_decoder_pack(const float* w1, const float* b1, const float* z, const float* w2,
                              const float* b2, const float* w3, const float* b3, int32_t s_dim,
                              int32_t z_dim, int32_t n_z, int32_t h1, int32_t h2, int32_t n_logvar,
                              void* packed, void* stream);
int klerg_target_decoder_pdf(const void* packed, int32_t s_dim, int32_t z_dim, int32_t n_z, int32_t h1,
                             int32_t h2, int32_t n_logvar, const float* samples, int64_t N,
                             const float* shift, float clamp_lo, float clamp_hi, float* p_out,
                             uint32_t* fault, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KLERG_B200_H */

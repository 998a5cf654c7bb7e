// C ABI of the fused KL-ergodic evals (kernels: klerg_fused.cuh, one translation unit per D: klerg_fused_d*.cu).
#include "klerg_fused.cuh"

namespace klerg {

FusedOptions g_fused_opt = {0, 0, 1, -1, 0};
EmuState g_emu = {};

template <int D> int launch_grad_d(EvalArgs& a, int64_t n_max, cudaStream_t stream);
template <int D> int launch_cost_d(EvalArgs& a, int64_t n_max, cudaStream_t stream);
extern template int launch_grad_d<1>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_grad_d<2>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_grad_d<3>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_grad_d<4>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_grad_d<5>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_grad_d<6>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_cost_d<1>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_cost_d<2>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_cost_d<3>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_cost_d<4>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_cost_d<5>(EvalArgs&, int64_t, cudaStream_t);
extern template int launch_cost_d<6>(EvalArgs&, int64_t, cudaStream_t);

const int* g_eval_gate = nullptr;

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

extern "C" int klerg_set_option(int key, int value) {
  switch (key) {
    case KLERG_OPT_EVAL_OVERLAP: g_fused_opt.overlap = value != 0; return 0;
    case KLERG_OPT_GRID_LIMIT: g_fused_opt.grid_limit = value > 0 ? value : 0; return 0;
    case KLERG_OPT_PDL: g_fused_opt.pdl = value != 0; return 0;
    case KLERG_OPT_EXACT_PAIRS: g_exact_pairs = value != 0; return 0;
    case KLERG_OPT_MIXED_WARPS: g_fused_opt.mixed_warps = (value == 12 || value == 16 || value == 1) ? value : 0; return 0;
    case KLERG_OPT_SATURATE_MILLI: g_saturate_milli = value > 0 ? value : 0; return 0;
    default: set_error("set_option: unknown key %d", key); return -1;
  }
}
extern "C" int klerg_get_option(int key) {
  switch (key) {
    case KLERG_OPT_EVAL_OVERLAP: return g_fused_opt.overlap;
    case KLERG_OPT_GRID_LIMIT: return g_fused_opt.grid_limit;
    case KLERG_OPT_PDL: return g_fused_opt.pdl;
    case KLERG_OPT_COOP_WITH_PDL: return g_fused_opt.coop_probe;
    case KLERG_OPT_EXACT_PAIRS: return g_exact_pairs;
    case KLERG_OPT_MIXED_WARPS: return g_fused_opt.mixed_warps;
    case KLERG_OPT_SATURATE_MILLI: return g_saturate_milli;
    default: return -1;
  }
}

extern "C" int klerg_emu_begin(void) {
  g_emu = EmuState{};
  g_emu.active = 1;
  return 0;
}
extern "C" int klerg_emu_launch(void* stream) {
  if (!g_emu.active) { set_error("emu_launch: no emulation in progress"); return -1; }
  g_emu.active = 0;
  if (!g_emu.have[0] || !g_emu.have[1] || !g_emu.launch) { set_error("emu_launch: both ranks must have recorded the same eval"); return -1; }
  if (2 * g_emu.nblk > sm_count()) { set_error("emu_launch: %d CTAs per rank do not fit twice on the device (set KLERG_OPT_GRID_LIMIT)", g_emu.nblk); return -1; }
  return g_emu.launch(g_emu.args[0], g_emu.args[1], g_emu.nblk, g_emu.nthreads, g_emu.smem, (cudaStream_t)stream);
}

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
extern "C" size_t klerg_debug_cta_stamps_offset(void) { return HEAD_COUNTERS + HEAD_MISC + HEAD_GRAD + FUSED_CTRL + MB_OFF_DBG; }

// single GPU: the mailbox is the region of the workspace reserved for it
static void finish_peers(EvalArgs& a, void* workspace) {
  if (a.peers.world <= 1) a.peers.mail[0] = ws_fused_mailbox(workspace);
  a.gate = g_eval_gate;
  a.independent = g_eval_gate ? 0 : g_fused_opt.overlap;
}

extern "C" int klerg_eval_gradient_targets(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn,
                                           const klerg_barrier_spec* bar, const klerg_peers* peers, const float* x0,
                                           const float* R0, const float* u, int64_t H, const float* packed, int64_t N,
                                           int64_t ld, const float* q_base, const float* p, int64_t K, int64_t p_stride,
                                           const double* p_stats, float floor, const float* Rinv_diag, float alpha,
                                           const float* ctrl_lo, const float* ctrl_hi, float* v_scratch, float* traj,
                                           double* totals, float* cost, float* dgdx, float* du, float* djdlam,
                                           float* u_star, double* kl_out, float* fault_out, void* workspace, void* stream) {
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
  a.ap.sat = 1e-3f * (float)g_saturate_milli;
  a.x0 = x0; a.R0 = R0; a.u = u; a.G = 1; a.H = (int)H; a.packed = packed; a.N = N; a.ld = ld; a.q_base = q_base; a.p = p;
  a.K = (int)K; a.p_stride = p_stride;
  a.p_stats = p_stats; a.floor = floor; a.v = v_scratch; a.ws = workspace; a.traj = traj; a.totals = totals; a.cost = cost;
  a.dgdx = dgdx; a.du = du; a.djdlam = djdlam; a.u_star = u_star; a.kl_out = kl_out; a.fault_out = fault_out;
  finish_peers(a, workspace);
  const int64_t n_max = (peers && peers->world > 1) ? peers->n_max : 0;
  switch (a.k.D) {
    case 1: return launch_grad_d<1>(a, n_max, (cudaStream_t)stream);
    case 2: return launch_grad_d<2>(a, n_max, (cudaStream_t)stream);
    case 3: return launch_grad_d<3>(a, n_max, (cudaStream_t)stream);
    case 4: return launch_grad_d<4>(a, n_max, (cudaStream_t)stream);
    case 5: return launch_grad_d<5>(a, n_max, (cudaStream_t)stream);
    case 6: return launch_grad_d<6>(a, n_max, (cudaStream_t)stream);
    default: set_error("eval_gradient: D=%d not instantiated (1..6)", a.k.D); return -2;
  }
}

extern "C" int klerg_eval_gradient(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                                   const klerg_peers* peers, const float* x0, const float* R0, const float* u, int64_t H,
                                   const float* packed, int64_t N, int64_t ld, const float* q_base, const float* p,
                                   const double* p_stats, float floor, const float* Rinv_diag, float alpha,
                                   const float* ctrl_lo, const float* ctrl_hi, float* v_scratch, float* traj,
                                   double* totals, float* cost, float* dgdx, float* du, float* djdlam, float* u_star,
                                   double* kl_out, float* fault_out, void* workspace, void* stream) {
  return klerg_eval_gradient_targets(k, dyn, bar, peers, x0, R0, u, H, packed, N, ld, q_base, p, 1, 0, p_stats, floor,
                                     Rinv_diag, alpha, ctrl_lo, ctrl_hi, v_scratch, traj, totals, cost, dgdx, du, djdlam,
                                     u_star, kl_out, fault_out, workspace, stream);
}

extern "C" int klerg_eval_costs(const klerg_kernel_spec* k, const klerg_dyn_spec* dyn, const klerg_barrier_spec* bar,
                                const klerg_peers* peers, const float* x0, const float* R0, const float* u, int64_t G,
                                int64_t H, const float* packed, int64_t N, int64_t ld, const float* q_base,
                                const float* p, const double* p_stats, float floor, float* v_scratch, float* traj,
                                double* totals, float* cost, float* fault_out, void* workspace, void* stream) {
  EvalArgs a{};
  if (!fill_common(a, k, dyn, bar, peers)) return -1;
  if (H < 1 || H > KLERG_MAX_H) { set_error("eval_costs: H out of range"); return -1; }
  if (G < 1 || G > FUSED_MAXG) { set_error("eval_costs: G must be in 1..%d", FUSED_MAXG); return -1; }
  if (N < 1 || ld < N || (ld & 3)) { set_error("eval_costs: bad sample sizes"); return -1; }
  if (!workspace || !v_scratch || !cost) { set_error("eval_costs: null output/workspace"); return -1; }
  if (((uintptr_t)packed | (uintptr_t)p | (uintptr_t)v_scratch) & 15) { set_error("eval_costs: packed, p and v_scratch must be 16-byte aligned"); return -1; }
  a.x0 = x0; a.R0 = R0; a.u = u; a.G = (int)G; a.H = (int)H; a.packed = packed; a.N = N; a.ld = ld; a.q_base = q_base;
  a.p = p; a.p_stats = p_stats; a.floor = floor; a.v = v_scratch; a.ws = workspace; a.traj = traj; a.totals = totals;
  a.cost = cost; a.K = 1; a.p_stride = 0; a.fault_out = fault_out;
  finish_peers(a, workspace);
  const int64_t n_max = (peers && peers->world > 1) ? peers->n_max : 0;
  switch (a.k.D) {
    case 1: return launch_cost_d<1>(a, n_max, (cudaStream_t)stream);
    case 2: return launch_cost_d<2>(a, n_max, (cudaStream_t)stream);
    case 3: return launch_cost_d<3>(a, n_max, (cudaStream_t)stream);
    case 4: return launch_cost_d<4>(a, n_max, (cudaStream_t)stream);
    case 5: return launch_cost_d<5>(a, n_max, (cudaStream_t)stream);
    case 6: return launch_cost_d<6>(a, n_max, (cudaStream_t)stream);
    default: set_error("eval_costs: D=%d not instantiated (1..6)", a.k.D); return -2;
  }
}

"""ctypes binding of libklerg_b200.so (see include/klerg_b200.h).

The library is the product: there is no CPU fallback.  Loading fails loudly if
the shared object is missing, and every wrapper raises ``RuntimeError`` when the
C side returns a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO_ROOT = os.path.dirname(PKG_ROOT)
# KLERG_VARIANT=_stamps selects the instrumented build (phase stamps in the fused evals, -DKLERG_STAMPS): a second
# library next to the product one, used by tools/ only
VARIANT = os.environ.get("KLERG_VARIANT", "")
LIB_PATH = os.path.join(PKG_ROOT, f"libklerg_b200{VARIANT}.so")
CSRC = os.path.join(PKG_ROOT, "csrc")
INCLUDE = os.path.join(REPO_ROOT, "include")

MAX_D, MAX_S, MAX_A, MAX_H = 8, 24, 8, 256
DYN_SINGLE, DYN_DOUBLE, DYN_SPEED, DYN_ROLL = 0, 1, 2, 3
RED_SUM, RED_MAX, RED_MIN = 0, 1, 2

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


class KernelSpec(C.Structure):
    _fields_ = [("D", C.c_int32), ("S", C.c_int32), ("explr", C.c_int32 * MAX_D),
                ("scale", C.c_float * MAX_D), ("nu", C.c_float)]


class DynSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("S", C.c_int32), ("A", C.c_int32), ("dt", C.c_float),
                ("rpw", C.c_int32 * 3), ("has_ang_map", C.c_int32),
                ("rot_lo", C.c_float * 3), ("rot_hi", C.c_float * 3),
                ("ang_lo", C.c_float * 3), ("ang_hi", C.c_float * 3)]


class Peers(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("mailbox", C.c_void_p * 8), ("n_max", C.c_int64)]


ABI_VERSION = 6  # mirrors klerg_abi_version() of the library built from this tree (include/klerg_b200.h)
OPT_EVAL_OVERLAP, OPT_GRID_LIMIT, OPT_PDL, OPT_COOP_WITH_PDL, OPT_EXACT_PAIRS, OPT_MIXED_WARPS = 1, 2, 3, 4, 5, 6
OPT_SATURATE_MILLI = 7


POLICY_LQR, POLICY_BARRIER_PUSH = 1, 2


class PolicySpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("use_u", C.c_int32), ("weight", C.c_float), ("K", C.c_float * (8 * 24))]


class TiltSpec(C.Structure):
    _fields_ = [("r_idx", C.c_int32), ("p_idx", C.c_int32), ("w_idx", C.c_int32), ("has_map", C.c_int32),
                ("w_lo", C.c_float), ("w_hi", C.c_float), ("tilt_lim", C.c_float), ("power", C.c_float), ("weight", C.c_float),
                ("rot_lo", C.c_float * 2), ("rot_hi", C.c_float * 2), ("ang_lo", C.c_float * 2), ("ang_hi", C.c_float * 2)]


class BarrierSpec(C.Structure):
    _fields_ = [("n", C.c_int32), ("lo", C.c_float * MAX_S), ("hi", C.c_float * MAX_S),
                ("weight", C.c_float * MAX_S), ("power", C.c_float * MAX_S)]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libklerg_b200.so (in-tree)."""
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [
        os.path.join(INCLUDE, "klerg_b200.h")]
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-DKLERG_STAMPS"] if (os.environ.get("KLERG_STAMPS") or VARIANT == "_stamps") else []
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + ["-I", INCLUDE]
    objdir = os.path.join(PKG_ROOT, "build" + VARIANT)
    os.makedirs(objdir, exist_ok=True)
    # one nvcc per translation unit, in parallel (the fused evals are the long pole), then one link
    procs = []
    for src in srcs:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + flags + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((obj, cmd, subprocess.Popen(cmd)))
    for obj, cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + [o for o, _, _ in procs]
    if verbose:
        print(" ".join(link))
    subprocess.run(link, check=True)
    return LIB_PATH


_lib = None

_P, _I64, _I32, _F = C.c_void_p, C.c_int64, C.c_int32, C.c_float
_KS, _DS, _BS = C.POINTER(KernelSpec), C.POINTER(DynSpec), C.POINTER(BarrierSpec)
_PS, _FP = C.POINTER(Peers), C.POINTER(C.c_float)

# name -> argtypes (restype int unless listed in _RESTYPES); mirrors include/klerg_b200.h
SIGNATURES = {
    "klerg_last_error": [],
    "klerg_abi_version": [],
    "klerg_launch_count": [],
    "klerg_peak_probe": [C.c_int, C.c_int, C.c_int, _P, _P],
    "klerg_device_info": [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "klerg_workspace_bytes": [_I64],
    "klerg_pack_samples": [_KS, _P, _I64, _P, _I64, _P],
    "klerg_footprint": [_KS, C.c_int, _P, _I64, _I64, _I64, _P, _I64, _I64, _P, _P, _I64, _P, _P, _P],
    "klerg_footprint_sum_max": [_KS, _P, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _P, _P],
    "klerg_footprint_tc_scratch_bytes": [_I64],
    "klerg_footprint_sum_max_tc": [_KS, _P, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _P, _P, _I64, _P],
    "klerg_psi_matrix": [_KS, _P, _I64, _P, _I64, _P, _P, _P],
    "klerg_vector_stats": [_P, _I64, _P, _P, _P],
    "klerg_renormalize": [_P, _I64, _F, _P, _P, _P],
    "klerg_renormalize_with_stats": [_P, _I64, _P, _F, _P, _P],
    "klerg_cost_norm": [_P, _I64, _P, _P],
    "klerg_kl_gradient": [_KS, _P, _I64, _P, _I64, _I64, _P, _P, _P, _P],
    "klerg_kl_gradient_fused": [_KS, _P, _I64, _P, _I64, _I64, _P, _P, C.c_int, _P, _F, _P, _P, _P, _P],
    "klerg_kl_cost_partial": [_P, _I64, _I64, _I64, _P, C.c_int, _P, _F, _P, _P, _P],
    "klerg_kl_cost_final": [_P, C.c_int, _I64, _P, _P, _P, _P],
    "klerg_target_stage1": [_P, _I32, _I64, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P, _P, _P, _P],
    "klerg_combine_blocks": [_P, C.c_int, C.c_int, C.POINTER(C.c_int), _P, _P],
    "klerg_target_exponent": [_P, _I64, _P, _P],
    "klerg_target_stage2": [C.c_int, _P, _I32, _I64, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P, _P, _P,
                            _P, _P, _P],
    "klerg_target_stage3": [_P, _I64, _P, C.c_int, _F, _F, _P, _P, _P, _P],
    "klerg_rollout": [_DS, _BS, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P],
    "klerg_barrier_eval": [_BS, _P, _I64, _I32, _P, _P, _P],
    "klerg_barrier_eval_ext": [_BS, _P, _P, C.POINTER(TiltSpec), _I64, _I32, _P, _P, _P, _P],
    "klerg_policy_rollout": [_DS, C.POINTER(PolicySpec), _P, _P, _P, _I64, _P, _P, _P],
    "klerg_adjoint_policy": [_DS, _I64, _P, _P, _P, _P, _P, C.POINTER(C.c_float), _F, C.POINTER(C.c_float),
                             C.POINTER(C.c_float), _P, _P, _P, _P],
    "klerg_adjoint": [_DS, _KS, _I64, _P, C.c_int, _P, _P, _P, _P, C.POINTER(C.c_float), _F,
                      C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P, _P, _P, _P],
    "klerg_adjoint_targets": [_DS, _KS, _I64, _I64, _P, C.c_int, _P, _P, _P, _P, C.POINTER(C.c_float), _F,
                              C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P, _P, _P, _P],
    "klerg_eval_costs_batch_scratch_bytes": [_I64, _I64, _I32],
    "klerg_eval_costs_batch": [_KS, _DS, _BS, _P, _P, _P, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P],
    "klerg_plan_scratch_bytes": [_I64, _I32, _I32, _I64],
    "klerg_plan_result_floats": [_I64, _I32, _I32],
    "klerg_plan_optimize": [_KS, _DS, _BS, _PS, _P, _P, _P, _I64, _P, _I64, _I64, _P, _P, _P, _F, _FP, _F, _FP, _FP,
                            _I32, _I32, _I32, _I32, _P, _P, _P, _P],
    "klerg_belief_scratch_bytes": [_I64, _I32],
    "klerg_belief_update": [_P, _I64, _I32, _P, _I32, C.c_double, C.c_double, _P, _P, _P, _P, _P, _P],
    "klerg_mt19937_uniform": [_P, _I32, _I32, _I64, _I32, _FP, _FP, _I64, _I64, _P, _P, _P],
    "klerg_gather_rows": [_P, _I32, _P, _I64, _P, _P],
    "klerg_mailbox_bytes": [],
    "klerg_mailbox_create": [C.POINTER(C.c_void_p), C.c_char_p],
    "klerg_mailbox_open": [C.c_char_p, C.POINTER(C.c_void_p)],
    "klerg_mailbox_close": [_P, C.c_int],
    "klerg_debug_stamps_offset": [],
    "klerg_debug_cta_stamps_offset": [],
    "klerg_fused_fault_offset": [],
    "klerg_set_option": [C.c_int, C.c_int],
    "klerg_get_option": [C.c_int],
    "klerg_emu_begin": [],
    "klerg_emu_launch": [_P],
    "klerg_eval_gradient": [_KS, _DS, _BS, _PS, _P, _P, _P, _I64, _P, _I64, _I64, _P, _P, _P, _F, _FP, _F, _FP, _FP,
                            _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "klerg_eval_gradient_targets": [_KS, _DS, _BS, _PS, _P, _P, _P, _I64, _P, _I64, _I64, _P, _P, _I64, _I64, _P, _F, _FP, _F,
                                    _FP, _FP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "klerg_kl_gradient_targets_scratch_bytes": [_I64, _I64],
    "klerg_kl_gradient_targets": [_KS, _P, _I64, _P, _I64, _I64, _P, _P, C.c_int, _P, _I64, _I64, _F, _P, _P, _P, _P, _P],
    "klerg_target_decoder_packed_bytes": [_I32, _I32, _I32, _I32, _I32, _I32],
    "klerg_target_decoder_pack": [_P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P],
    "klerg_target_decoder_pdf": [_P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _I64, _P, _F, _F, _P, _P, _P],
    "klerg_eval_costs": [_KS, _DS, _BS, _PS, _P, _P, _P, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _F, _P, _P, _P, _P,
                         _P, _P, _P],
}
_RESTYPES = {"klerg_last_error": C.c_char_p, "klerg_target_decoder_packed_bytes": C.c_size_t, "klerg_kl_gradient_targets_scratch_bytes": C.c_size_t, "klerg_workspace_bytes": C.c_size_t, "klerg_mailbox_bytes": C.c_size_t, "klerg_debug_stamps_offset": C.c_size_t, "klerg_debug_cta_stamps_offset": C.c_size_t, "klerg_fused_fault_offset": C.c_size_t,
             "klerg_launch_count": C.c_longlong, "klerg_eval_costs_batch_scratch_bytes": C.c_size_t, "klerg_belief_scratch_bytes": C.c_size_t, "klerg_plan_scratch_bytes": C.c_size_t,
             "klerg_plan_result_floats": C.c_int64, "klerg_footprint_tc_scratch_bytes": C.c_int64}


def load():
    """dlopen the library and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the KL-ergodic path.")
    lib = C.CDLL(LIB_PATH)
    lib.klerg_abi_version.restype = C.c_int
    got = lib.klerg_abi_version()
    if got != ABI_VERSION:  # a stale .so would have its arguments misbound without any error
        raise RuntimeError(f"{LIB_PATH} has ABI version {got}, this package binds version {ABI_VERSION}: rebuild it with "
                           "`python -c 'import __graft_entry__ as g; g.build()'`")
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if os.environ.get("KLERG_PDL") == "0":  # A/B switch: plain cooperative launches of the fused evals
        lib.klerg_set_option(OPT_PDL, 0)
    if os.environ.get("KLERG_MIXED_WARPS") in ("1", "12", "16"):  # A/B switch: gradient schedules for D >= 5
        lib.klerg_set_option(OPT_MIXED_WARPS, int(os.environ["KLERG_MIXED_WARPS"]))
    if os.environ.get("KLERG_EXACT_PAIRS") == "1":  # A/B switch: difference form of the squared distance everywhere
        lib.klerg_set_option(OPT_EXACT_PAIRS, 1)
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("control_torch (B200 build) needs a CUDA device; there is no CPU fallback")


def check(status, what):
    if status != 0:
        msg = load().klerg_last_error()
        raise RuntimeError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


def raw_stream():
    """cudaStream_t of torch's current stream as an int (the fast path: torch.cuda.current_stream() builds a Python
    Stream object on every call, ~15 us, which adds up over the ~30 library calls of a planner step)."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def stream_ptr():
    return C.c_void_p(raw_stream())


def farr(vals, n=None):
    vals = [float(v) for v in vals]
    n = n or len(vals)
    return (C.c_float * n)(*vals)


def kernel_spec(D, S, explr, scale, nu=1.0):
    k = KernelSpec()
    k.D, k.S, k.nu = int(D), int(S), float(nu)
    for d in range(int(D)):
        k.explr[d] = int(explr[d])
        k.scale[d] = float(scale[d])
    return k


def dyn_spec(kind, S, A, dt, rpw=(0, 0, 0), ang_map=None):
    d = DynSpec()
    d.kind, d.S, d.A, d.dt = int(kind), int(S), int(A), float(dt)
    for i in range(3):
        d.rpw[i] = int(rpw[i])
    d.has_ang_map = 0
    if ang_map is not None:
        rot, ang = ang_map  # [3,2] each
        d.has_ang_map = 1
        for i in range(3):
            d.rot_lo[i], d.rot_hi[i] = float(rot[i][0]), float(rot[i][1])
            d.ang_lo[i], d.ang_hi[i] = float(ang[i][0]), float(ang[i][1])
    return d


def policy_spec(kind, use_u=0, weight=0.0, K=None):
    p = PolicySpec()
    p.kind, p.use_u, p.weight = int(kind), int(bool(use_u)), float(weight)
    if K is not None:
        flat = [float(v) for row in K for v in row]
        if len(flat) > MAX_A * MAX_S:
            raise ValueError("policy gain larger than KLERG_MAX_A x KLERG_MAX_S")
        for i, v in enumerate(flat):
            p.K[i] = v
    return p


def barrier_spec(lo, hi, weight, power):
    b = BarrierSpec()
    b.n = len(lo)
    if b.n > MAX_S:
        raise ValueError("barrier has more rows than KLERG_MAX_S")
    for i in range(b.n):
        b.lo[i], b.hi[i], b.weight[i], b.power[i] = float(lo[i]), float(hi[i]), float(weight[i]), float(power[i])
    return b

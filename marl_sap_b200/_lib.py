"""ctypes binding of libmarl_sap_b200.so (the C ABI declared in include/marl_sap_b200.h).

There is NO CPU fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

PKG = os.path.dirname(os.path.abspath(__file__))
# SAP_ABLATE=1: the profiling build with the timing-ablation hooks (SAP_ABLATE=1 python -m marl_sap_b200._build)
LIB_PATH = os.path.join(PKG, "libmarl_sap_b200_ablate.so" if os.environ.get("SAP_ABLATE") == "1" else "libmarl_sap_b200.so")

SAP_F32, SAP_F16, SAP_I64, SAP_I16, SAP_I32, SAP_U8 = range(6)

_TORCH2SAP = {
    torch.float32: SAP_F32, torch.float16: SAP_F16, torch.int64: SAP_I64, torch.int16: SAP_I16,
    torch.int32: SAP_I32, torch.uint8: SAP_U8, torch.bool: SAP_U8,
}


def sap_dtype(dt: torch.dtype) -> int:
    try:
        return _TORCH2SAP[dt]
    except KeyError:
        raise TypeError(f"marl_sap_b200: dtype {dt} has no device representation in the C ABI") from None


class SapEnvDims(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("B", "n", "m", "T", "L", "M", "N", "shared_planes")]


class SapField(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("env_stride", C.c_int64), ("t_stride", C.c_int64),
                ("dtype", C.c_int32), ("reserved", C.c_int32)]


VIEW_FIELDS = ("obs", "rewards", "actions", "actions_onehot", "terminated", "filled", "prev_assigns", "beta",
               "avail_actions", "agent_in")


class SapBatchView(C.Structure):
    _fields_ = [(k, SapField) for k in VIEW_FIELDS]


class SapSelectArgs(C.Structure):
    """Selector side of ``sap_rollout_step`` (selection + env step in one launch)."""
    _fields_ = [("q", C.c_void_p), ("eps_dev", C.c_void_p), ("episode_ctr", C.c_void_p), ("u_explore", C.c_void_p),
                ("u_action", C.c_void_p), ("seed", C.c_uint64), ("eps", C.c_float), ("reserved", C.c_int32)]


# name -> (restype, argtypes); mirrors include/marl_sap_b200.h one to one
_P, _I32, _I64, _U64, _F32, _F64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
_DIMS, _VIEW = C.POINTER(SapEnvDims), C.POINTER(SapBatchView)
SIGNATURES = {
    "sap_abi_version": (C.c_int, []),
    "sap_last_error": (C.c_char_p, []),
    "sap_benefit_ingest": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P]),
    "sap_benefit_upload_host": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "sap_benefit_stats": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P]),
    "sap_benefit_generate": (C.c_int, [_P, _I32, _I32, _I32, _I32, _F32, _F32, _U64, _U64, _P]),
    "sap_real_reset": (C.c_int, [_DIMS, _P, _P, _P, _P, _P, _P, _VIEW, _P, _P, _P]),
    "sap_real_step": (C.c_int, [_DIMS, _P, _P, _P, _P, _F64, _P, _P, _P, _P, _P, _VIEW, _P, _P, _P]),
    "sap_real_scratch_doubles": (C.c_int64, [_DIMS]),
    "sap_proximities_fov": (C.c_int, [_P, _P, _I32, _I32, _I32, _F64, _F64, _P, _P, _P]),
    "sap_real_reset_ex": (C.c_int, [_DIMS, _P, _P, _P, _P, _P, _P, _VIEW, _P, _P, _P, _I32, _P]),
    "sap_real_step_ex": (C.c_int, [_DIMS, _P, _P, _P, _P, _F64, _P, _P, _P, _P, _P, _VIEW, _P, _P, _P, _I32, _P]),
    "sap_power_pre": (C.c_int, [_DIMS, _P, _P, _P, _P, _P, _P, _P]),
    "sap_power_post": (C.c_int, [_DIMS, _P, _P, _P, _VIEW, C.POINTER(SapField), _I32, _I32, _P]),
    "sap_interference_rewards": (C.c_int, [_DIMS, _P, _P, _P, _P, _I32, _F64, _P, _P, _P, _P, _P, C.POINTER(SapField), _P]),
    "sap_mock_reset": (C.c_int, [_DIMS, _P, _P, _P, _P, _P, _VIEW, _P]),
    "sap_mock_step": (C.c_int, [_DIMS, _P, _P, _F64, _P, _P, _P, _P, _P, _VIEW, _P]),
    "sap_select_epsilon_greedy": (C.c_int, [_P, _P, _I32, _I32, _I32, _F32, _P, _U64, _P, _P, _P, _P, _P, _P]),
    "sap_select_filtered_epsilon_greedy": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _F32, _P, _U64, _P, _P, _P, _P,
                                                     _P, _P, _P]),
    "sap_topm_from_beta": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "sap_buffer_insert": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _I64, _P]),
    "sap_buffer_gather": (C.c_int, [_P, _P, _P, _I64, _I64, _P]),
    "sap_onehot": (C.c_int, [_P, _I32, _P, _I32, _I64, _I32, _P]),
    "sap_real_beta_window": (C.c_int, [_DIMS, _P, _P, _P, _I32, _P]),
    "sap_bias_act": (C.c_int, [_P, _P, _I64, _I32, _I32, _P]),
    "sap_split_bias_act": (C.c_int, [_P, _I32, _F32, _P, _P, _I64, _I32, _I32, _P]),
    "sap_real_agent_in_f16_ok": (C.c_int, [_DIMS]),
    "sap_real_obs_ahead_ok": (C.c_int, [_DIMS]),
    "sap_real_obs_ahead": (C.c_int, [_DIMS, _P, _P, _P, _VIEW, _P, _P]),
    "sap_real_step_after_obs": (C.c_int, [_DIMS, _P, _P, _P, _F64, _P, _P, _P, _P, _P, _VIEW, _P, _P]),
    "sap_rollout_step_ok": (C.c_int, [_DIMS]),
    "sap_rollout_step": (C.c_int, [C.POINTER(SapSelectArgs), _DIMS, _P, _P, _P, _F64, _P, _P, _P, _P, _P, _VIEW, _P, _P, _P]),
    "sap_real_beta_rows": (C.c_int, [_DIMS, _P, _P, _P, _I32, _I32, _P, _I32, _P]),
    "sap_real_select_kernel": (C.c_int32, [_I32]),
    "sap_lsa_maximize": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _P, _P]),
    "sap_sample_categorical": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"marl_sap_b200: {LIB_PATH} not found. Build it with `python -m marl_sap_b200._build` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    ver = lib.sap_abi_version()
    if ver != 1:
        raise RuntimeError(f"marl_sap_b200: ABI version mismatch (library {ver}, binding 1)")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sap_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"marl_sap_b200.{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device (or host) address of a tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"marl_sap_b200: `{name}` must live on a CUDA device; there is no CPU path")


def field_of(t, B_T_leading: bool = True) -> SapField:
    """SapField for a [B, T+1, ...] tensor whose trailing dims are contiguous."""
    f = SapField()
    if t is None:
        return f
    assert t.dim() >= 2
    inner = 1
    for s, st in zip(reversed(t.shape[2:]), reversed(t.stride()[2:])):
        assert s == 1 or st == inner, "episode-batch field must be contiguous in its trailing dims"
        inner *= s
    f.ptr, f.env_stride, f.t_stride, f.dtype = t.data_ptr(), t.stride(0), t.stride(1), sap_dtype(t.dtype)
    return f


(REAL_PATH_AUTO, REAL_PATH_GENERIC, REAL_PATH_LARGE_KEYED, REAL_PATH_LARGE_EXACT, REAL_PATH_FAST_GEN1,
 REAL_PATH_FAST_RUNTIME_SHAPE) = range(6)


class select_real_kernel:
    """Context manager around ``sap_real_select_kernel``: run the real-env step on one named kernel path (tests and
    profiles run every path on the same shape; production code leaves the choice to the library)."""

    def __init__(self, which):
        self.which = int(which)

    def __enter__(self):
        self.prev = load().sap_real_select_kernel(self.which)
        if self.prev < 0:
            raise ValueError(f"unknown real-env kernel path {self.which}")
        return self

    def __exit__(self, *exc):
        load().sap_real_select_kernel(self.prev)
        return False

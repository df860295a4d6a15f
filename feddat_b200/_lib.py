"""ctypes binding of ``libfeddat_sm100.so`` (C ABI declared in ``include/feddat_b200.h``).

The library is the product: if it cannot be loaded the import fails loudly -- there is no
PyTorch/CPU fallback behind these entry points.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_uint32, c_void_p
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libfeddat_sm100.so"
_DBG_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libfeddat_sm100_dbg.so"
_lib = None
_dbg_lib = None
ABI_VERSION = 5      # feddat_abi_version(); bumped whenever include/feddat_b200.h changes

# every symbol include/feddat_b200.h declares (tests check the built library exports all of them)
EXPORTED_SYMBOLS = (
    "feddat_last_error",
    "feddat_abi_version",
    "feddat_dat_fwd",
    "feddat_dat_bwd_dgrad",
    "feddat_dat_bwd_wgrad",
    "feddat_dat_wgrad_workspace_bytes",
    "feddat_dat_fwd_grouped",
    "feddat_dat_bwd_dgrad_grouped",
    "feddat_dat_bwd_wgrad_grouped",
    "feddat_pack_weights",
    "feddat_pack_weights_batched",
    "feddat_mkd_loss",
    "feddat_mkd_ce_loss",
    "feddat_fedavg",
    "feddat_ln_fwd",
    "feddat_ln_bwd",
    "feddat_gelu_fwd",
    "feddat_gelu_bwd",
    "feddat_mlp_fc1_gelu_fwd",
    "feddat_mlp_fc2_dgelu_bwd",
    "feddat_attn_fwd",
    "feddat_attn_bwd",
    "feddat_attn_bwd_workspace_bytes",
    "feddat_patchify",
    "feddat_adamw_step",
)
# include/feddat_b200_debug.h: only in the -DFEDDAT_DEBUG twin (libfeddat_sm100_dbg.so), tests / scripts
DEBUG_SYMBOLS = (
    "feddat_probe_gemm",
    "feddat_debug_set_trace",
    "feddat_probe_l2bw",
    "feddat_debug_force_fused_fwd",
    "feddat_probe_pair",
    "feddat_probe_ingest",
    "feddat_probe_tilecopy",
)


class FeddatError(RuntimeError):
    pass


class DatGroup(ctypes.Structure):
    """``FeddatDatGroup`` of include/feddat_b200.h."""
    _fields_ = [("X", c_void_p), ("Res", c_void_p), ("Y", c_void_p), ("dY", c_void_p), ("dX", c_void_p),
                ("Wd_cat", c_void_p), ("bd_cat", c_void_p), ("Wu_cat", c_void_p), ("bu_cat", c_void_p),
                ("WuT_cat", c_void_p), ("WdT_cat", c_void_p), ("H_out", c_void_p), ("H_in", c_void_p),
                ("H_t", c_void_p), ("dP_t", c_void_p), ("ld_t", c_int), ("r_lo", c_int), ("r_hi", c_int),
                ("M", c_int64), ("r_total", c_int), ("branch_scale", c_float), ("add_dy", c_int)]


class WgradGroup(ctypes.Structure):
    """``FeddatWgradGroup`` of include/feddat_b200.h."""
    _fields_ = [("X", c_void_p), ("dY", c_void_p), ("H_t", c_void_p), ("dP_t", c_void_p),
                ("dWu", c_void_p), ("dbu", c_void_p), ("dWd", c_void_p), ("dbd", c_void_p),
                ("M", c_int64), ("r_t", c_int), ("ld_ht", c_int), ("ld_dwu", c_int), ("branch_scale", c_float)]


class AdamwTensor(ctypes.Structure):
    """FeddatAdamwTensor (include/feddat_b200.h)."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("step", c_void_p), ("lr", c_void_p), ("weight_decay", c_float), ("numel", c_int64)]


class PackJob(ctypes.Structure):
    """``FeddatPackJob`` of include/feddat_b200.h."""
    _fields_ = [("down_w", c_void_p * 2), ("down_b", c_void_p * 2), ("up_w", c_void_p * 2), ("bu_src", c_void_p * 2),
                ("n_branch", c_int), ("r", c_int), ("ld_up", c_int),
                ("Wd_cat", c_void_p), ("WdT_cat", c_void_p), ("Wu_cat", c_void_p), ("WuT_cat", c_void_p),
                ("bd_cat", c_void_p), ("bu_cat", c_void_p)]


def lib_path() -> Path:
    return _LIB_PATH


def _open(path: Path, debug: bool) -> ctypes.CDLL:
    if not path.exists():
        if os.environ.get("FEDDAT_NO_AUTOBUILD"):
            raise FeddatError(f"{path} is missing; run `python -m feddat_b200.build`")
        from .build import build_library
        build_library(debug=debug)          # lock file + atomic rename: safe under torchrun
    return ctypes.CDLL(str(path))


def load_debug() -> ctypes.CDLL:
    """The -DFEDDAT_DEBUG twin: every product entry point (with trace hooks compiled in) plus the probes and
    debug switches of include/feddat_b200_debug.h.  Tests and scripts only; the package never calls this."""
    global _dbg_lib
    if _dbg_lib is None:
        _dbg_lib = _bind(_open(_DBG_LIB_PATH, True), debug=True)
    return _dbg_lib


def load() -> ctypes.CDLL:
    global _lib
    if os.environ.get("FEDDAT_DEBUG_LIB"):       # scripts/trace_kernel.py: route ops.* through the traced twin
        return load_debug()
    if _lib is None:
        _lib = _bind(_open(_LIB_PATH, False), debug=False)
    return _lib


def _bind(lib: ctypes.CDLL, debug: bool) -> ctypes.CDLL:
    lib.feddat_last_error.restype = c_char_p
    lib.feddat_last_error.argtypes = []
    lib.feddat_abi_version.restype = c_int
    lib.feddat_abi_version.argtypes = []

    lib.feddat_dat_fwd.restype = c_int
    lib.feddat_dat_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_int, c_int, c_void_p]
    lib.feddat_dat_bwd_dgrad.restype = c_int
    lib.feddat_dat_bwd_dgrad.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64,
                                         c_int, c_int, c_float, c_int, c_int, c_int, c_void_p]
    lib.feddat_dat_bwd_wgrad.restype = c_int
    lib.feddat_dat_bwd_wgrad.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                         c_float, c_int, c_void_p, c_size_t, c_void_p]
    lib.feddat_dat_wgrad_workspace_bytes.restype = c_size_t
    lib.feddat_dat_wgrad_workspace_bytes.argtypes = []
    lib.feddat_dat_fwd_grouped.restype = c_int
    lib.feddat_dat_fwd_grouped.argtypes = [POINTER(DatGroup), c_int, c_int, c_int, c_int, c_void_p]
    lib.feddat_dat_bwd_dgrad_grouped.restype = c_int
    lib.feddat_dat_bwd_dgrad_grouped.argtypes = [POINTER(DatGroup), c_int, c_int, c_int, c_int, c_void_p]
    lib.feddat_dat_bwd_wgrad_grouped.restype = c_int
    lib.feddat_dat_bwd_wgrad_grouped.argtypes = [POINTER(WgradGroup), c_int, c_int, c_int, c_void_p, c_size_t,
                                                 c_void_p]
    lib.feddat_pack_weights.restype = c_int
    lib.feddat_pack_weights.argtypes = [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                        POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.feddat_pack_weights_batched.restype = c_int
    lib.feddat_pack_weights_batched.argtypes = [POINTER(PackJob), c_int, c_int, c_void_p]
    lib.feddat_mkd_loss.restype = c_int
    lib.feddat_mkd_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                    c_float, c_float, c_float, c_float, c_int64, c_void_p, c_void_p]
    lib.feddat_mkd_ce_loss.restype = c_int
    lib.feddat_mkd_ce_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                       c_int, c_int, c_float, c_float, c_float, c_int, c_void_p, c_void_p]
    lib.feddat_fedavg.restype = c_int
    lib.feddat_fedavg.argtypes = [POINTER(c_void_p), POINTER(c_float), c_int, c_float, c_void_p,
                                  c_int64, c_void_p]
    lib.feddat_ln_fwd.restype = c_int
    lib.feddat_ln_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int64, c_int, c_float, c_int, c_void_p]
    lib.feddat_ln_bwd.restype = c_int
    lib.feddat_ln_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_int, c_int, c_void_p]
    lib.feddat_gelu_fwd.restype = c_int
    lib.feddat_gelu_fwd.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p]
    lib.feddat_gelu_bwd.restype = c_int
    lib.feddat_gelu_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]
    lib.feddat_mlp_fc1_gelu_fwd.restype = c_int
    lib.feddat_mlp_fc1_gelu_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                            c_void_p]
    lib.feddat_mlp_fc2_dgelu_bwd.restype = c_int
    lib.feddat_mlp_fc2_dgelu_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]
    lib.feddat_attn_fwd.restype = c_int
    lib.feddat_attn_fwd.argtypes = [c_void_p] * 5 + [c_int] * 4 + [c_int64] * 4 + [c_float, c_int, c_void_p]
    lib.feddat_attn_bwd.restype = c_int
    lib.feddat_attn_bwd.argtypes = [c_void_p] * 9 + [c_int] * 4 + [c_int64] * 8 + [c_float, c_void_p, ctypes.c_size_t, c_int,
                                    c_void_p]
    lib.feddat_adamw_step.restype = c_int
    lib.feddat_adamw_step.argtypes = [POINTER(AdamwTensor), c_int, c_float, c_float, c_float, c_void_p]
    lib.feddat_patchify.restype = c_int
    lib.feddat_patchify.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.feddat_attn_bwd_workspace_bytes.restype = ctypes.c_size_t
    lib.feddat_attn_bwd_workspace_bytes.argtypes = [c_int, c_int]
    if not debug:
        return lib
    lib.feddat_probe_gemm.restype = c_int
    lib.feddat_probe_gemm.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      POINTER(c_uint32), c_void_p]
    lib.feddat_probe_l2bw.restype = c_int
    lib.feddat_probe_l2bw.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.feddat_debug_force_fused_fwd.restype = c_int
    lib.feddat_debug_force_fused_fwd.argtypes = [c_int]
    lib.feddat_probe_tilecopy.restype = c_int
    lib.feddat_probe_tilecopy.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]
    lib.feddat_probe_ingest.restype = c_int
    lib.feddat_probe_ingest.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p]
    lib.feddat_probe_pair.restype = c_int
    lib.feddat_probe_pair.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                      c_void_p]
    lib.feddat_debug_set_trace.restype = c_int
    lib.feddat_debug_set_trace.argtypes = [c_void_p]
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().feddat_last_error().decode("utf-8", "replace")
        raise FeddatError(f"{what or 'libfeddat_sm100'} failed (code {rc}): {msg}")


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (or None -> NULL)."""
    return c_void_p(0 if t is None else t.data_ptr())


def stream_ptr() -> c_void_p:
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)

"""Torch-tensor front end of the C ABI (``include/feddat_b200.h``): argument checking, output
allocation through PyTorch's caching allocator, launch on the current stream.  Every function here
ends in a hand-written sm_100a kernel; none has a PyTorch fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib

ACT_RELU, ACT_GELU = 0, 1
_ACT = {"relu": ACT_RELU, "gelu": ACT_GELU}
DTYPE_BF16 = 0
MAX_R_TOTAL = 256      # widest bottleneck one dat_fwd / dat_bwd_dgrad launch covers
MAX_R_WGRAD = 128      # widest trainable slice one dat_bwd_wgrad launch covers

launch_count = 0       # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def act_code(act) -> int:
    return _ACT[act] if isinstance(act, str) else int(act)


def _check_act2d(t: torch.Tensor, name: str, d: int = 768) -> None:
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.shape[1] == d and t.is_contiguous()):
        raise _lib.FeddatError(
            f"{name}: expected a contiguous CUDA bf16 [M, {d}] tensor, got {tuple(t.shape)} {t.dtype} "
            f"{t.device} contiguous={t.is_contiguous()} (no CPU / fp32 fallback exists)")


@dataclass
class PackedWeights:
    """bf16 operands of one adapter site in one mode (see feddat_pack_weights)."""
    wd: torch.Tensor     # [R, d]
    wdT: torch.Tensor    # [d, R]
    wu: torch.Tensor     # [d, R]
    wuT: torch.Tensor    # [R, d]
    bd: torch.Tensor     # [R] fp32
    bu: torch.Tensor     # [d] fp32
    r: int               # per-branch rank
    n_branch: int

    @property
    def r_total(self) -> int:
        return self.r * self.n_branch


def pack_weights(branches: Sequence[Sequence[torch.Tensor]], need_bwd: bool = True) -> PackedWeights:
    """branches: [(down_w [r,d], down_b [r], up_w [d,r], up_b [d]), ...] fp32 CUDA tensors (1 or 2)."""
    lib = _lib.load()
    nb = len(branches)
    dw0 = branches[0][0]
    r, d = dw0.shape
    dev = dw0.device
    for b in branches:
        for t, shp in zip(b, ((r, d), (r,), (d, r), (d,))):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shp):
                raise _lib.FeddatError(f"pack_weights: expected contiguous CUDA fp32 {shp}, got "
                                       f"{tuple(t.shape)} {t.dtype} {t.device}")
    R = nb * r
    bf = dict(device=dev, dtype=torch.bfloat16)
    wd = torch.empty(R, d, **bf)
    wu = torch.empty(d, R, **bf)
    wdT = torch.empty(d, R, **bf) if need_bwd else None
    wuT = torch.empty(R, d, **bf) if need_bwd else None
    bd = torch.empty(R, device=dev, dtype=torch.float32)
    bu = torch.empty(d, device=dev, dtype=torch.float32)
    arr = ctypes.c_void_p * nb
    tabs = [arr(*[b[i].data_ptr() for b in branches]) for i in range(4)]
    rc = lib.feddat_pack_weights(tabs[0], tabs[1], tabs[2], tabs[3], nb, r, d, _lib.ptr(wd),
                                 _lib.ptr(wdT), _lib.ptr(wu), _lib.ptr(wuT), _lib.ptr(bd),
                                 _lib.ptr(bu), _lib.stream_ptr())
    _lib.check(rc, "feddat_pack_weights")
    _count()
    return PackedWeights(wd, wdT, wu, wuT, bd, bu, r, nb)


def dat_forward(x: torch.Tensor, res: torch.Tensor, w: PackedWeights, scale: float, act=ACT_RELU,
                out: Optional[torch.Tensor] = None, save_hidden: bool = False):
    """Y = res + scale * (act(x Wd^T + bd) Wu^T + bu)   (adapter.py:124-163).  With ``save_hidden`` returns
    (Y, H) where H [M, r_total] bf16 is the hidden, for a backward that does not recompute it."""
    lib = _lib.load()
    _check_act2d(x, "dat_forward x")
    _check_act2d(res, "dat_forward res")
    if w.r_total > MAX_R_TOTAL:
        raise _lib.FeddatError(f"dat_forward: r_total={w.r_total} > {MAX_R_TOTAL}; the Adapter module "
                               "splits such bottlenecks into several launches")
    y = out if out is not None else torch.empty_like(x)
    h = torch.empty(x.shape[0], w.r_total, device=x.device, dtype=torch.bfloat16) if save_hidden else None
    rc = lib.feddat_dat_fwd(_lib.ptr(x), _lib.ptr(res), _lib.ptr(y), _lib.ptr(w.wd), _lib.ptr(w.bd),
                            _lib.ptr(w.wu), _lib.ptr(w.bu), _lib.ptr(h), x.shape[0], x.shape[1], w.r_total,
                            float(scale), act_code(act), DTYPE_BF16, _lib.stream_ptr())
    _lib.check(rc, "feddat_dat_fwd")
    _count()
    return (y, h) if save_hidden else y


def dat_backward(x: Optional[torch.Tensor], dy: torch.Tensor, w: PackedWeights, scale: float, act=ACT_RELU,
                 train_slice: Optional[tuple] = None, need_dx: bool = True, add_dy: bool = True,
                 hidden: Optional[torch.Tensor] = None, dx_out: Optional[torch.Tensor] = None):
    """Backward of dat_forward.  Returns (dx | None, grads | None) where grads =
    (d_down_w [rt,d], d_down_b [rt], d_up_w [d,rt], d_up_b [d]) in fp32 for the trainable slice
    ``train_slice = (r_lo, r_hi)`` of the concatenated bottleneck.  ``hidden`` = the H saved by
    ``dat_forward(save_hidden=True)`` (ReLU): the dgrad kernel then skips the recompute of x Wd^T
    (``x`` is still needed by the weight-gradient kernel when something trains)."""
    lib = _lib.load()
    _check_act2d(dy, "dat_backward dy")
    if x is not None:
        _check_act2d(x, "dat_backward x")
    if w.wdT is None:
        raise _lib.FeddatError("dat_backward: weights were packed with need_bwd=False")
    saved = hidden is not None
    if saved and act_code(act) != ACT_RELU:
        raise _lib.FeddatError("dat_backward: a saved hidden determines act' only for ReLU")
    if x is None and (not saved or train_slice is not None):
        raise _lib.FeddatError("dat_backward: x is required (recompute mode, or weight gradients)")
    M, d = dy.shape
    R = w.r_total
    dev = dy.device
    if dx_out is not None:
        _check_act2d(dx_out, "dat_backward dx_out")
    dx = (dx_out if dx_out is not None else torch.empty_like(dy)) if need_dx else None
    h_t = dp_t = None
    rt = 0
    if train_slice is not None:
        r_lo, r_hi = train_slice
        rt = r_hi - r_lo
        if saved:
            # full-width scratch: only the trainable columns are written / read (row stride R, like H)
            dp_full = torch.empty(M, R, device=dev, dtype=torch.bfloat16)
            dp_t = dp_full[:, r_lo:r_hi]
            h_t = hidden[:, r_lo:r_hi]
            ld_t = R
        else:
            h_t = torch.empty(M, rt, device=dev, dtype=torch.bfloat16)
            dp_t = torch.empty(M, rt, device=dev, dtype=torch.bfloat16)
            ld_t = rt
    else:
        r_lo = r_hi = 0
        ld_t = 0
    if dx is None and dp_t is None:
        return None, None
    grads = None
    if train_slice is not None:
        f32 = dict(device=dev, dtype=torch.float32)
        g = torch.zeros(2 * d * rt + rt + d, **f32)          # one memset for all four gradients
    rc = lib.feddat_dat_bwd_dgrad(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dx), _lib.ptr(w.wd), _lib.ptr(w.bd),
                                  _lib.ptr(w.wuT), _lib.ptr(w.wdT), _lib.ptr(hidden),
                                  None if saved else _lib.ptr(h_t), _lib.ptr(dp_t), ld_t, r_lo,
                                  r_hi, M, d, R, float(scale), act_code(act), int(add_dy),
                                  DTYPE_BF16, _lib.stream_ptr())
    _lib.check(rc, "feddat_dat_bwd_dgrad")
    _count()
    if train_slice is not None:
        d_down_w = g[: rt * d].view(rt, d)
        d_up_w = g[rt * d: 2 * rt * d].view(d, rt)
        d_down_b = g[2 * rt * d: 2 * rt * d + rt]
        d_up_b = g[2 * rt * d + rt:]
        for j0 in range(0, rt, MAX_R_WGRAD):
            w_ = min(MAX_R_WGRAD, rt - j0)
            rc = lib.feddat_dat_bwd_wgrad(
                _lib.ptr(x), _lib.ptr(dy), ctypes.c_void_p(h_t.data_ptr() + 2 * j0),
                ctypes.c_void_p(dp_t.data_ptr() + 2 * j0), ctypes.c_void_p(d_up_w.data_ptr() + 4 * j0),
                _lib.ptr(d_up_b) if j0 == 0 else None, ctypes.c_void_p(d_down_w.data_ptr() + 4 * j0 * d),
                ctypes.c_void_p(d_down_b.data_ptr() + 4 * j0), M, d, w_, ld_t, rt, float(scale), DTYPE_BF16,
                _lib.stream_ptr())
            _lib.check(rc, "feddat_dat_bwd_wgrad")
            _count()
        grads = (d_down_w, d_down_b, d_up_w, d_up_b)
    return dx, grads


def mkd_loss(logits: torch.Tensor, teacher: torch.Tensor, target: Optional[torch.Tensor], temp: float,
             kl_weight: float = 0.5, task_weight: float = 0.5, need_grad: bool = True):
    """Fused (task + kl_loss)/2 head (task_trainer.py:299-301, 506-516) for 2-D [rows, C] fp32
    logits.  Returns (loss3, dlogits): loss3 = device tensor [total, kl, task]."""
    lib = _lib.load()
    for t, n in ((logits, "logits"), (teacher, "teacher"), (target, "target")):
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape == logits.shape):
            raise _lib.FeddatError(f"mkd_loss: {n} must be a contiguous CUDA fp32 tensor shaped like logits")
    C = logits.shape[-1]
    rows = logits.numel() // C
    # reference kl_loss: softmax over dim=-1 when C > 3000, over dim=1 otherwise (task_trainer.py:507-512);
    # both coincide with "last dim" for the 2-D ViLT logits and the 3-D ALBEF logits it is used on
    if logits.dim() > 2 and C <= 3000:
        raise _lib.FeddatError("mkd_loss: >2-D logits with C <= 3000 would softmax over dim=1 in the "
                               "reference; that layout is not implemented")
    batchmean_div = logits.shape[0]
    task_scale = 1.0 / logits.shape[0] if target is not None else 0.0
    loss3 = torch.empty(3, device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits) if need_grad else None
    rc = lib.feddat_mkd_loss(_lib.ptr(logits), _lib.ptr(teacher), _lib.ptr(target), _lib.ptr(loss3),
                             _lib.ptr(dlogits), rows, C, float(temp), float(kl_weight), float(task_weight),
                             float(task_scale), batchmean_div, _lib.stream_ptr())
    _lib.check(rc, "feddat_mkd_loss")
    _count(2)  # memset + kernel
    return loss3, dlogits


def fedavg(client_bufs: Sequence[torch.Tensor], nums: Sequence[float], out: torch.Tensor,
           total: float = 0.0) -> torch.Tensor:
    """out = sum_c client_c * num_c / total, reference operation order (main.py:57-64);
    total defaults to sum(nums) (pass the global total for a per-rank partial sum)."""
    lib = _lib.load()
    n = out.numel()
    for t in list(client_bufs) + [out]:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n):
            raise _lib.FeddatError("fedavg: all buffers must be contiguous CUDA fp32 of equal length")
    nc = len(client_bufs)
    ptrs = (ctypes.c_void_p * nc)(*[t.data_ptr() for t in client_bufs])
    w = (ctypes.c_float * nc)(*[float(v) for v in nums])
    rc = lib.feddat_fedavg(ptrs, w, nc, float(total), _lib.ptr(out), n, _lib.stream_ptr())
    _lib.check(rc, "feddat_fedavg")
    _count()
    return out


def layer_norm_fwd(x: torch.Tensor, res: Optional[torch.Tensor], weight: torch.Tensor, bias: torch.Tensor,
                   eps: float, bias2: Optional[torch.Tensor] = None):
    """(y, s, mean, rstd, s2): s = bf16(x + res) (s is x itself when res is None), y = LayerNorm(s) over the
    last dimension (768), statistics in fp32; s2 = bf16(s + bias2) when ``bias2`` is given, else None
    (feddat_ln_fwd)."""
    lib = _lib.load()
    _check_act2d(x, "layer_norm x")
    if res is not None:
        _check_act2d(res, "layer_norm res")
    for t, n in ((weight, "weight"), (bias, "bias"), (bias2, "bias2")):
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() == x.shape[1]):
            raise _lib.FeddatError(f"layer_norm {n}: expected a contiguous CUDA bf16 [{x.shape[1]}] tensor")
    M, d = x.shape
    y = torch.empty_like(x)
    s = torch.empty_like(x) if res is not None else None
    s2 = torch.empty_like(x) if bias2 is not None else None
    stats = torch.empty(2, M, device=x.device, dtype=torch.float32)
    rc = lib.feddat_ln_fwd(_lib.ptr(x), _lib.ptr(res), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(y), _lib.ptr(s),
                           _lib.ptr(bias2), _lib.ptr(s2), _lib.ptr(stats[0]), _lib.ptr(stats[1]), M, d, float(eps),
                           DTYPE_BF16, _lib.stream_ptr())
    _lib.check(rc, "feddat_ln_fwd")
    _count()
    return y, (s if s is not None else x), stats[0], stats[1], s2


def layer_norm_bwd(dy: torch.Tensor, dsum: Optional[torch.Tensor], s: torch.Tensor, weight: torch.Tensor,
                   mean: torch.Tensor, rstd: torch.Tensor) -> torch.Tensor:
    """dL/ds of layer_norm_fwd for FROZEN affine parameters, plus ``dsum`` (the gradient reaching ``s``
    through its other consumer, the residual stream) when given (feddat_ln_bwd)."""
    lib = _lib.load()
    _check_act2d(dy, "layer_norm_bwd dy")
    _check_act2d(s, "layer_norm_bwd s")
    if dsum is not None:
        _check_act2d(dsum, "layer_norm_bwd dsum")
    dx = torch.empty_like(s)
    rc = lib.feddat_ln_bwd(_lib.ptr(dy), _lib.ptr(dsum), _lib.ptr(s), _lib.ptr(weight), _lib.ptr(mean),
                           _lib.ptr(rstd), _lib.ptr(dx), s.shape[0], s.shape[1], DTYPE_BF16, _lib.stream_ptr())
    _lib.check(rc, "feddat_ln_bwd")
    _count()
    return dx


def gelu_fwd(x: torch.Tensor) -> torch.Tensor:
    """Exact-erf GELU of a contiguous CUDA bf16 tensor (feddat_gelu_fwd)."""
    lib = _lib.load()
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() % 8 == 0):
        raise _lib.FeddatError("gelu: expected a contiguous CUDA bf16 tensor with a multiple of 8 elements")
    y = torch.empty_like(x)
    _lib.check(lib.feddat_gelu_fwd(_lib.ptr(x), _lib.ptr(y), x.numel(), DTYPE_BF16, _lib.stream_ptr()), "feddat_gelu_fwd")
    _count()
    return y


def gelu_bwd(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    for t in (dy, x):
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape == x.shape):
            raise _lib.FeddatError("gelu_bwd: expected contiguous CUDA bf16 tensors of one shape")
    dx = torch.empty_like(x)
    _lib.check(lib.feddat_gelu_bwd(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dx), x.numel(), DTYPE_BF16, _lib.stream_ptr()),
               "feddat_gelu_bwd")
    _count()
    return dx

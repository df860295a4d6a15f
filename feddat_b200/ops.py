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


@dataclass
class PackSpec:
    """One packing job (``FeddatPackJob``): per-branch fp32 CUDA tensors.  ``down_w[b]`` [r, d] and
    ``down_b[b]`` [r] contiguous (a row slice of the master is fine), ``up_w[b]`` [d, r] with unit column
    stride (a column slice view of the master is fine: its row stride is passed along), ``bu_src`` the up
    biases summed into ``bu`` (entries may be None)."""
    down_w: Sequence[torch.Tensor]
    down_b: Sequence[torch.Tensor]
    up_w: Sequence[torch.Tensor]
    bu_src: Sequence[Optional[torch.Tensor]]
    need_bwd: bool = True


def alloc_packed(r: int, nb: int, d: int, device, need_bwd: bool = True) -> PackedWeights:
    R = nb * r
    bf = dict(device=device, dtype=torch.bfloat16)
    return PackedWeights(torch.empty(R, d, **bf), torch.empty(d, R, **bf) if need_bwd else None,
                         torch.empty(d, R, **bf), torch.empty(R, d, **bf) if need_bwd else None,
                         torch.empty(R, device=device, dtype=torch.float32),
                         torch.empty(d, device=device, dtype=torch.float32), r, nb)


def pack_weights_batched(specs: Sequence[PackSpec], outs: Optional[Sequence[PackedWeights]] = None):
    """All ``specs`` in ONE launch (feddat_pack_weights_batched; 24 jobs per launch).  ``outs``: persistent
    operand sets to overwrite (allocated when None).  Returns the list of PackedWeights."""
    lib = _lib.load()
    if not specs:
        return []
    d = specs[0].down_w[0].shape[1]
    dev = specs[0].down_w[0].device
    res = []
    jobs = (_lib.PackJob * len(specs))()
    for i, sp in enumerate(specs):
        nb = len(sp.down_w)
        r = sp.down_w[0].shape[0]
        for b in range(nb):
            dw, db, uw = sp.down_w[b], sp.down_b[b], sp.up_w[b]
            ok = (dw.is_cuda and dw.dtype == torch.float32 and dw.is_contiguous() and tuple(dw.shape) == (r, d)
                  and db.is_cuda and db.dtype == torch.float32 and db.is_contiguous() and tuple(db.shape) == (r,)
                  and uw.is_cuda and uw.dtype == torch.float32 and tuple(uw.shape) == (d, r) and uw.stride(1) == 1)
            if not ok:
                raise _lib.FeddatError(f"pack_weights: job {i} branch {b}: expected fp32 CUDA down_w [{r}, {d}] / down_b "
                                       f"[{r}] contiguous and up_w [{d}, {r}] with unit column stride, got "
                                       f"{tuple(dw.shape)} {tuple(db.shape)} {tuple(uw.shape)} strides {uw.stride()}")
        for t in sp.bu_src:
            if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (d,)):
                raise _lib.FeddatError(f"pack_weights: job {i}: up biases must be contiguous fp32 CUDA [{d}]")
        out = outs[i] if outs is not None else alloc_packed(r, nb, d, dev, sp.need_bwd)
        if out.r != r or out.n_branch != nb:
            raise _lib.FeddatError("pack_weights: persistent operand set has another shape")
        j = jobs[i]
        for b in range(nb):
            j.down_w[b], j.down_b[b], j.up_w[b] = sp.down_w[b].data_ptr(), sp.down_b[b].data_ptr(), sp.up_w[b].data_ptr()
        for b, t in enumerate(list(sp.bu_src)[:2]):
            j.bu_src[b] = None if t is None else t.data_ptr()
        j.n_branch, j.r, j.ld_up = nb, r, sp.up_w[0].stride(0)
        if any(u.stride(0) != j.ld_up for u in sp.up_w):
            raise _lib.FeddatError("pack_weights: the branches of one job must share the up_w row stride")
        j.Wd_cat, j.Wu_cat = out.wd.data_ptr(), out.wu.data_ptr()
        j.WdT_cat = None if out.wdT is None or not sp.need_bwd else out.wdT.data_ptr()
        j.WuT_cat = None if out.wuT is None or not sp.need_bwd else out.wuT.data_ptr()
        j.bd_cat, j.bu_cat = out.bd.data_ptr(), out.bu.data_ptr()
        res.append(out)
    rc = lib.feddat_pack_weights_batched(jobs, len(specs), d, _lib.stream_ptr())
    _lib.check(rc, "feddat_pack_weights_batched")
    _count(-(-len(specs) // 24))
    return res


def pack_weights(branches: Sequence[Sequence[torch.Tensor]], need_bwd: bool = True) -> PackedWeights:
    """branches: [(down_w [r,d], down_b [r], up_w [d,r], up_b [d]), ...] fp32 CUDA tensors (1 or 2)."""
    spec = PackSpec([b[0] for b in branches], [b[1] for b in branches], [b[2] for b in branches],
                    [b[3] for b in branches], need_bwd)
    for b in branches:
        if not b[2].is_contiguous():
            raise _lib.FeddatError("pack_weights: expected contiguous CUDA fp32 tensors")
    return pack_weights_batched([spec])[0]


def _fwd_group(x, res, w, scale, out, save_hidden):
    _check_act2d(x, "dat_forward x")
    _check_act2d(res, "dat_forward res")
    if w.r_total > MAX_R_TOTAL:
        raise _lib.FeddatError(f"dat_forward: r_total={w.r_total} > {MAX_R_TOTAL}; the Adapter module "
                               "splits such bottlenecks into several launches")
    y = out if out is not None else torch.empty_like(x)
    h = torch.empty(x.shape[0], w.r_total, device=x.device, dtype=torch.bfloat16) if save_hidden else None
    g = _lib.DatGroup()
    g.X, g.Res, g.Y = x.data_ptr(), res.data_ptr(), y.data_ptr()
    g.Wd_cat, g.bd_cat, g.Wu_cat, g.bu_cat = w.wd.data_ptr(), w.bd.data_ptr(), w.wu.data_ptr(), w.bu.data_ptr()
    g.H_out = None if h is None else h.data_ptr()
    g.M, g.r_total, g.branch_scale = x.shape[0], w.r_total, float(scale)
    return g, y, h


def dat_forward_grouped(groups: Sequence[dict], act=ACT_RELU):
    """Up to two independent row groups in ONE launch (feddat_dat_fwd_grouped): each group is a dict with the
    arguments of ``dat_forward`` (x, res, w, scale, out=None, save_hidden=False).  Returns [(y, h | None)]."""
    lib = _lib.load()
    arr = (_lib.DatGroup * len(groups))()
    outs = []
    for i, gr in enumerate(groups):
        g, y, h = _fwd_group(gr["x"], gr["res"], gr["w"], gr["scale"], gr.get("out"), gr.get("save_hidden", False))
        arr[i] = g
        outs.append((y, h))
    rc = lib.feddat_dat_fwd_grouped(arr, len(groups), groups[0]["x"].shape[1], act_code(act), DTYPE_BF16,
                                    _lib.stream_ptr())
    _lib.check(rc, "feddat_dat_fwd_grouped")
    _count()
    return outs


def dat_forward(x: torch.Tensor, res: torch.Tensor, w: PackedWeights, scale: float, act=ACT_RELU,
                out: Optional[torch.Tensor] = None, save_hidden: bool = False):
    """Y = res + scale * (act(x Wd^T + bd) Wu^T + bu)   (adapter.py:124-163).  With ``save_hidden`` returns
    (Y, H) where H [M, r_total] bf16 is the hidden, for a backward that does not recompute it."""
    (y, h), = dat_forward_grouped([dict(x=x, res=res, w=w, scale=scale, out=out, save_hidden=save_hidden)], act)
    return (y, h) if save_hidden else y


_wgrad_ws = {}


def wgrad_workspace(device: torch.device) -> torch.Tensor:
    """Workspace of the deterministic weight-gradient reduction: one per device, zero-initialised once (the
    kernel leaves its counters at zero).  The process model is one client training per GPU at a time
    (SURVEY.md section 8e): weight-gradient launches of one device are stream-ordered and share it; launches
    that could run CONCURRENTLY on one device would need a workspace each (C ABI: feddat_dat_bwd_wgrad)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    ws = _wgrad_ws.get(key)
    if ws is None:
        nbytes = int(_lib.load().feddat_dat_wgrad_workspace_bytes())
        if nbytes == 0:
            raise _lib.FeddatError("feddat_dat_wgrad_workspace_bytes() failed (no sm_100 device?)")
        if torch.cuda.is_current_stream_capturing():
            raise _lib.FeddatError("the weight-gradient workspace must exist before CUDA-graph capture: run one "
                                   "eager backward on this stream first (GraphedTrainStep's warm-up does)")
        ws = _wgrad_ws[key] = torch.zeros(nbytes // 4 + 4, device=device, dtype=torch.float32)
    return ws


MAX_WGRAD_GROUPS = 24        # groups of one feddat_dat_bwd_wgrad_grouped launch


def _launch_wgrad(wg, device) -> None:
    lib = _lib.load()
    ws = wgrad_workspace(device)
    for i in range(0, len(wg), MAX_WGRAD_GROUPS):
        chunk = wg[i:i + MAX_WGRAD_GROUPS]
        arr = (_lib.WgradGroup * len(chunk))(*chunk)
        rc = lib.feddat_dat_bwd_wgrad_grouped(arr, len(chunk), 768, DTYPE_BF16, _lib.ptr(ws), ws.numel() * 4,
                                              _lib.stream_ptr())
        _lib.check(rc, "feddat_dat_bwd_wgrad_grouped")
        _count()


class DeferredWgrad:
    """Weight gradients of a whole backward pass in ONE launch.

    The parameter gradients of an adapter site are not needed before the optimizer step, while its data gradient
    is on the critical path of back-propagation.  Inside ``with ops.deferred_wgrad() as q:`` every
    ``dat_backward_grouped`` call launches only its data-gradient kernel and queues its weight-gradient groups
    (keeping X, dY, the saved hidden and dP alive); ``q.flush()`` -- called by the ``with`` exit -- launches them
    together (24 groups per launch: 12 ViLT sites x {gating rows, adapter_1 rows} = 144 CTAs, each contracting over
    ALL rows of its (group, 128-column chunk): no row splits, no partial tiles, no second reduction stage, one
    launch instead of twelve) and then hands the gradients to the parameters registered with ``assign`` exactly as
    autograd's AccumulateGrad would (``p.grad = g`` or ``p.grad += g``).  Process-global on purpose: autograd runs
    CUDA backward nodes on its own thread."""

    def __init__(self):
        self.groups, self.keep, self.pending = [], [], []
        self.device = None

    def add(self, wgroups, keep, device) -> None:
        self.groups += wgroups
        self.keep.append(keep)
        self.device = device

    def assign(self, param: torch.Tensor, grad_view: torch.Tensor) -> None:
        self.pending.append((param, grad_view))

    def flush(self) -> None:
        if self.groups:
            # every CTA of a launch walks ALL rows of its group: launch groups of similar size together (ALBEF: 9 232-row
            # ViT sites next to 400-row text sites would all wait for the longest)
            self.groups.sort(key=lambda q: -int(q.M))
            _launch_wgrad(self.groups, self.device)
        with torch.no_grad():
            for p, g in self.pending:
                g = g.contiguous()
                if p.grad is None:
                    p.grad = g
                else:
                    p.grad.add_(g)
        self.groups, self.keep, self.pending = [], [], []


_deferred: Optional[DeferredWgrad] = None


def deferred_queue() -> Optional[DeferredWgrad]:
    return _deferred


class deferred_wgrad:
    """Context manager: see DeferredWgrad."""

    def __enter__(self) -> DeferredWgrad:
        global _deferred
        if _deferred is not None:
            raise _lib.FeddatError("deferred_wgrad() does not nest")
        _deferred = DeferredWgrad()
        return _deferred

    def __exit__(self, exc_type, exc, tb):
        global _deferred
        q, _deferred = _deferred, None
        if exc_type is None:
            q.flush()
        return False


def dat_backward_grouped(groups: Sequence[dict], act=ACT_RELU, allow_defer: bool = False):
    """Backward of up to two row groups: ONE data-gradient launch (feddat_dat_bwd_dgrad_grouped) and ONE
    weight-gradient launch (feddat_dat_bwd_wgrad_grouped) for all of them -- or, inside ``deferred_wgrad()``, the
    weight-gradient groups of a caller that passes ``allow_defer`` (one that registers the returned gradient
    tensors with the queue instead of reading them) are queued and those tensors are filled at the flush.  Each group is a dict with the
    arguments of ``dat_backward`` (x, dy, w, scale, train_slice=None, need_dx=True, add_dy=True, hidden=None,
    dx_out=None).  Returns [(dx | None, grads | None)]."""
    lib = _lib.load()
    arr = (_lib.DatGroup * len(groups))()
    info = []
    n_live = 0
    d = groups[0]["dy"].shape[1]
    for gr in groups:
        x, dy, w, scale = gr.get("x"), gr["dy"], gr["w"], gr["scale"]
        train_slice, need_dx = gr.get("train_slice"), gr.get("need_dx", True)
        add_dy, hidden, dx_out = gr.get("add_dy", True), gr.get("hidden"), gr.get("dx_out")
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
        M = dy.shape[0]
        R = w.r_total
        dev = dy.device
        if dx_out is not None:
            _check_act2d(dx_out, "dat_backward dx_out")
        dx = (dx_out if dx_out is not None else torch.empty_like(dy)) if need_dx else None
        h_t = dp_t = None
        rt = r_lo = r_hi = ld_t = 0
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
        if dx is None and dp_t is None:
            info.append(None)
            continue
        g = arr[n_live]
        g.X, g.dY, g.dX = _lib.ptr(x), dy.data_ptr(), _lib.ptr(dx)
        g.Wd_cat, g.bd_cat, g.WuT_cat, g.WdT_cat = w.wd.data_ptr(), w.bd.data_ptr(), w.wuT.data_ptr(), w.wdT.data_ptr()
        g.H_in = _lib.ptr(hidden)
        g.H_t = None if saved else _lib.ptr(h_t)
        g.dP_t = _lib.ptr(dp_t)
        g.ld_t, g.r_lo, g.r_hi = ld_t, r_lo, r_hi
        g.M, g.r_total, g.branch_scale, g.add_dy = M, R, float(scale), int(add_dy)
        n_live += 1
        info.append(dict(x=x, dy=dy, dx=dx, h_t=h_t, dp_t=dp_t, ld_t=ld_t, rt=rt, scale=float(scale), M=M, dev=dev))
    if n_live:
        rc = lib.feddat_dat_bwd_dgrad_grouped(arr, n_live, d, act_code(act), DTYPE_BF16, _lib.stream_ptr())
        _lib.check(rc, "feddat_dat_bwd_dgrad_grouped")
        _count()
    # weight gradients: every <= 128-wide slice of every group is one "wgrad group"; two per launch
    wg, results = [], []
    for it in info:
        if it is None:
            results.append((None, None))
            continue
        grads = None
        if it["rt"] and it["M"] > 0:
            rt = it["rt"]
            g = torch.empty(2 * d * rt + rt + d, device=it["dev"], dtype=torch.float32)   # the kernel overwrites
            d_down_w = g[: rt * d].view(rt, d)
            d_up_w = g[rt * d: 2 * rt * d].view(d, rt)
            d_down_b = g[2 * rt * d: 2 * rt * d + rt]
            d_up_b = g[2 * rt * d + rt:]
            for j0 in range(0, rt, MAX_R_WGRAD):
                w_ = min(MAX_R_WGRAD, rt - j0)
                q = _lib.WgradGroup()
                q.X, q.dY = it["x"].data_ptr(), it["dy"].data_ptr()
                q.H_t, q.dP_t = it["h_t"].data_ptr() + 2 * j0, it["dp_t"].data_ptr() + 2 * j0
                q.dWu, q.dbu = d_up_w.data_ptr() + 4 * j0, (d_up_b.data_ptr() if j0 == 0 else None)
                q.dWd, q.dbd = d_down_w.data_ptr() + 4 * j0 * d, d_down_b.data_ptr() + 4 * j0
                q.M, q.r_t, q.ld_ht, q.ld_dwu, q.branch_scale = it["M"], w_, it["ld_t"], rt, it["scale"]
                wg.append(q)
            grads = (d_down_w, d_down_b, d_up_w, d_up_b)
        elif it["rt"]:
            g = torch.zeros(2 * d * it["rt"] + it["rt"] + d, device=it["dev"], dtype=torch.float32)
            rt = it["rt"]
            grads = (g[: rt * d].view(rt, d), g[2 * rt * d: 2 * rt * d + rt], g[rt * d: 2 * rt * d].view(d, rt),
                     g[2 * rt * d + rt:])
        results.append((it["dx"], grads))
    if wg:
        dev = groups[0]["dy"].device
        if _deferred is not None and allow_defer:
            wgrad_workspace(dev)       # exists before any capture
            _deferred.add(wg, (info, results, [gr.get("hidden") for gr in groups]), dev)
        else:
            for i in range(0, len(wg), 2):      # the groups of one site: two per launch (row splits fill the SMs)
                _launch_wgrad(wg[i:i + 2], dev)
    return results


def dat_backward(x: Optional[torch.Tensor], dy: torch.Tensor, w: PackedWeights, scale: float, act=ACT_RELU,
                 train_slice: Optional[tuple] = None, need_dx: bool = True, add_dy: bool = True,
                 hidden: Optional[torch.Tensor] = None, dx_out: Optional[torch.Tensor] = None,
                 allow_defer: bool = False):
    """Backward of dat_forward.  Returns (dx | None, grads | None) where grads =
    (d_down_w [rt,d], d_down_b [rt], d_up_w [d,rt], d_up_b [d]) in fp32 for the trainable slice
    ``train_slice = (r_lo, r_hi)`` of the concatenated bottleneck.  ``hidden`` = the H saved by
    ``dat_forward(save_hidden=True)`` (ReLU): the dgrad kernel then skips the recompute of x Wd^T
    (``x`` is still needed by the weight-gradient kernel when something trains)."""
    (dx, grads), = dat_backward_grouped([dict(x=x, dy=dy, w=w, scale=scale, train_slice=train_slice, need_dx=need_dx,
                                              add_dy=add_dy, hidden=hidden, dx_out=dx_out)], act, allow_defer=allow_defer)
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
    row_ws = torch.empty(max(rows, 1), 2, device=logits.device, dtype=torch.float32)
    rc = lib.feddat_mkd_loss(_lib.ptr(logits), _lib.ptr(teacher), _lib.ptr(target), _lib.ptr(loss3),
                             _lib.ptr(dlogits), rows, C, float(temp), float(kl_weight), float(task_weight),
                             float(task_scale), batchmean_div, _lib.ptr(row_ws), _lib.stream_ptr())
    _lib.check(rc, "feddat_mkd_loss")
    _count(2)  # row kernel + fixed-order final sum
    return loss3, dlogits


DTYPE_F32 = 1


def mkd_ce_loss(scores: torch.Tensor, teacher: torch.Tensor, labels: torch.Tensor, seq_weight: torch.Tensor,
                temp: float, kl_weight: float = 0.5, task_weight: float = 0.5, need_grad: bool = True):
    """Fused MKD head of the ALBEF path (feddat_mkd_ce_loss): KL(T) between the decoder logits and the teacher's +
    the weighted answer cross-entropy, value and d/dscores in one pass.  ``scores``: the UNSHIFTED prediction scores
    [n_seq, La, C] (bf16 or fp32, contiguous); ``teacher``: [n_seq, La, C] or the shifted [n_seq, La - 1, C], same
    dtype (a [:, :-1] VIEW of an unshifted tensor is taken as its base); ``labels`` [n_seq, La] int64 with -100 =
    ignore; ``seq_weight`` [n_seq] fp32 = answer weight / image batch.  Returns (loss3 [total, kl, task], dscores)."""
    lib = _lib.load()
    if not (scores.is_cuda and scores.dim() == 3 and scores.is_contiguous()
            and scores.dtype in (torch.bfloat16, torch.float32)):
        raise _lib.FeddatError("mkd_ce_loss: scores must be a contiguous CUDA [n_seq, La, C] bf16 / fp32 tensor")
    n_seq, La, C = scores.shape
    if teacher.dtype != scores.dtype:
        teacher = teacher.to(scores.dtype)
    if (not teacher.is_contiguous() and teacher.dim() == 3 and teacher.shape == (n_seq, La - 1, C)
            and teacher.stride() == (La * C, C, 1)):
        La_t = La                       # a [:, :-1] view of an unshifted tensor: address it through its row stride
    else:
        teacher = teacher.contiguous()
        La_t = teacher.shape[1]
    if teacher.shape[0] != n_seq or teacher.shape[2] != C or La_t not in (La, La - 1):
        raise _lib.FeddatError(f"mkd_ce_loss: teacher {tuple(teacher.shape)} does not match scores {tuple(scores.shape)}")
    if not (labels.is_cuda and labels.dtype == torch.int64 and tuple(labels.shape) == (n_seq, La)):
        raise _lib.FeddatError("mkd_ce_loss: labels must be a CUDA int64 [n_seq, La] tensor")
    labels = labels.contiguous()
    seq_weight = seq_weight.to(torch.float32).contiguous()
    if tuple(seq_weight.shape) != (n_seq,):
        raise _lib.FeddatError("mkd_ce_loss: seq_weight must be [n_seq]")
    loss3 = torch.empty(3, device=scores.device, dtype=torch.float32)
    dscores = torch.empty_like(scores) if need_grad else None
    row_ws = torch.empty(max(n_seq * La, 1), 2, device=scores.device, dtype=torch.float32)
    rc = lib.feddat_mkd_ce_loss(_lib.ptr(scores), _lib.ptr(teacher), _lib.ptr(labels), _lib.ptr(seq_weight),
                                _lib.ptr(loss3), _lib.ptr(dscores), n_seq, La, La_t, C, float(temp), float(kl_weight),
                                float(task_weight), DTYPE_BF16 if scores.dtype == torch.bfloat16 else DTYPE_F32,
                                _lib.ptr(row_ws), _lib.stream_ptr())
    _lib.check(rc, "feddat_mkd_ce_loss")
    _count(2)
    return loss3, dscores


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


def _check_bf16_2d(t: torch.Tensor, name: str) -> None:
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.is_contiguous() and t.data_ptr() % 16 == 0):
        raise _lib.FeddatError(f"{name}: expected a contiguous 16-byte aligned CUDA bf16 matrix, got {tuple(t.shape)} "
                               f"{t.dtype} {t.device}")


def mlp_fc1_gelu(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor):
    """(pre, act) = (a w^T + bias, gelu(pre)) with the exact GELU in the GEMM's epilogue (feddat_mlp_fc1_gelu_fwd).
    a [M, K], w [N, K] (nn.Linear weight): bf16; bias [N]: fp32 (a widened copy of the frozen bias)."""
    lib = _lib.load()
    _check_bf16_2d(a, "mlp_fc1_gelu a")
    _check_bf16_2d(w, "mlp_fc1_gelu w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or bias.shape != (N,) or bias.dtype != torch.float32 or not bias.is_contiguous():
        raise _lib.FeddatError("mlp_fc1_gelu: shapes / dtypes of w, bias do not match a")
    pre = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    act = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    rc = lib.feddat_mlp_fc1_gelu_fwd(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(pre), _lib.ptr(act), M, N, K,
                                     DTYPE_BF16, _lib.stream_ptr())
    _lib.check(rc, "feddat_mlp_fc1_gelu_fwd")
    _count()
    return pre, act


def mlp_fc2_dgelu(dy: torch.Tensor, w2t: torch.Tensor, pre: torch.Tensor) -> torch.Tensor:
    """dpre = (dy w2t^T) * gelu'(pre) (feddat_mlp_fc2_dgelu_bwd).  dy [M, K], w2t [N, K] = the transpose of the
    second dense layer's weight [K, N], pre [M, N]: bf16."""
    lib = _lib.load()
    _check_bf16_2d(dy, "mlp_fc2_dgelu dy")
    _check_bf16_2d(w2t, "mlp_fc2_dgelu w2t")
    _check_bf16_2d(pre, "mlp_fc2_dgelu pre")
    M, K = dy.shape
    N = w2t.shape[0]
    if w2t.shape[1] != K or tuple(pre.shape) != (M, N):
        raise _lib.FeddatError("mlp_fc2_dgelu: shapes of w2t / pre do not match dy")
    out = torch.empty_like(pre)
    rc = lib.feddat_mlp_fc2_dgelu_bwd(_lib.ptr(dy), _lib.ptr(w2t), _lib.ptr(pre), _lib.ptr(out), M, N, K, DTYPE_BF16,
                                      _lib.stream_ptr())
    _lib.check(rc, "feddat_mlp_fc2_dgelu_bwd")
    _count()
    return out


def _token_view(t: torch.Tensor, what: str):
    """(B, S, H, D, ld) of a bf16 [B, S, H, D] tensor whose (h, d) are contiguous and whose batch stride is S x the
    token stride -- a projection output [B * S, H * D] or a column slice of a wider one, viewed per head."""
    if t.dtype != torch.bfloat16 or not t.is_cuda or t.dim() != 4:
        raise _lib.FeddatError(f"{what}: expected a 4-D bf16 CUDA tensor [B, S, H, D]")
    B, S, H, D = t.shape
    sb, ss, sh, sd = t.stride()
    if sd != 1 or sh != D or (B > 1 and sb != S * ss) or ss < H * D:
        raise _lib.FeddatError(f"{what}: strides {t.stride()} are not a [B * S, ld] token layout")
    return B, S, H, D, ss


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float):
    """(o, lse) = softmax(q k^T scale) v per (batch, head) (feddat_attn_fwd).  q, k, v: [B, S, H, 64] bf16 views with
    (h, d) contiguous (token stride free); o: [B, S, H, 64] contiguous; lse: [B, H, S] fp32."""
    lib = _lib.load()
    B, S, H, D, ldq = _token_view(q, "attn_fwd q")
    _, _, _, _, ldk = _token_view(k, "attn_fwd k")
    _, _, _, _, ldv = _token_view(v, "attn_fwd v")
    if k.shape != q.shape or v.shape != q.shape:
        raise _lib.FeddatError("attn_fwd: q, k, v shapes differ")
    o = torch.empty(B, S, H, D, device=q.device, dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device=q.device, dtype=torch.float32)
    rc = lib.feddat_attn_fwd(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), _lib.ptr(lse), B, S, H, D, ldq, ldk, ldv,
                             H * D, float(scale), DTYPE_BF16, _lib.stream_ptr())
    _lib.check(rc, "feddat_attn_fwd")
    _count()
    return o, lse


ATTN_MAX_S_FWD = 256
ATTN_MAX_S_BWD = 192


def attn_bwd(do: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, lse: torch.Tensor,
             scale: float, packed: bool = False):
    """(dq, dk, dv) of ``attn_fwd`` (feddat_attn_bwd: a row-statistics pass, then one fused launch).  The three gradients
    are the column slices of ONE [B, S, 3, H, 64] tensor (token stride 3 * H * 64), returned whole with ``packed``: a
    fused q/k/v projection's backward can consume it as it is."""
    lib = _lib.load()
    B, S, H, D, ldq = _token_view(q, "attn_bwd q")
    _, _, _, _, ldk = _token_view(k, "attn_bwd k")
    _, _, _, _, ldv = _token_view(v, "attn_bwd v")
    _, _, _, _, ldo = _token_view(o, "attn_bwd o")
    _, _, _, _, lddo = _token_view(do, "attn_bwd do")
    if lse.dtype != torch.float32 or tuple(lse.shape) != (B, H, S) or not lse.is_contiguous():
        raise _lib.FeddatError("attn_bwd: lse must be a contiguous fp32 [B, H, S] tensor")
    dqkv = torch.empty(B, S, 3, H, D, device=q.device, dtype=torch.bfloat16)
    dq, dk, dv = dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]
    ld = 3 * H * D
    ws = torch.empty(B * H * 384, device=q.device, dtype=torch.float32)     # per-row statistics of the pre-pass
    rc = lib.feddat_attn_bwd(_lib.ptr(do), _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), _lib.ptr(lse), _lib.ptr(dq),
                             _lib.ptr(dk), _lib.ptr(dv), B, S, H, D, lddo, ldq, ldk, ldv, ldo, ld, ld, ld, float(scale),
                             _lib.ptr(ws), ws.numel() * 4, DTYPE_BF16, _lib.stream_ptr())
    _count()
    _lib.check(rc, "feddat_attn_bwd")
    _count()
    return dqkv if packed else (dq, dk, dv)


def patchify(pixel_values: torch.Tensor, patch: int) -> torch.Tensor:
    """[B, C, H, W] fp32 / bf16 -> [B (H // patch) (W // patch), C patch patch] bf16 rows of a stride = kernel convolution
    (feddat_patchify): cut and cast in one pass.  Trailing rows / columns that do not fill a patch are dropped, as the
    convolution drops them."""
    lib = _lib.load()
    if not (pixel_values.is_cuda and pixel_values.dim() == 4 and pixel_values.is_contiguous()
            and pixel_values.dtype in (torch.float32, torch.bfloat16)):
        raise _lib.FeddatError("patchify: expected a contiguous CUDA fp32 / bf16 [B, C, H, W] tensor")
    B, C, H, W = pixel_values.shape
    out = torch.empty(B * (H // patch) * (W // patch), C * patch * patch, device=pixel_values.device, dtype=torch.bfloat16)
    rc = lib.feddat_patchify(_lib.ptr(pixel_values), _lib.ptr(out), B, C, H, W, patch,
                             DTYPE_F32 if pixel_values.dtype == torch.float32 else DTYPE_BF16, _lib.stream_ptr())
    _lib.check(rc, "feddat_patchify")
    _count()
    return out


def adamw_step(tensors, n: int, beta1: float, beta2: float, eps: float) -> None:
    """One AdamW step over ``n`` FeddatAdamwTensor entries (a ctypes array built by train.fused_adamw.FusedAdamW)."""
    lib = _lib.load()
    rc = lib.feddat_adamw_step(tensors, n, float(beta1), float(beta2), float(eps), _lib.stream_ptr())
    _lib.check(rc, "feddat_adamw_step")
    _count(2 * ((n + 47) // 48))

"""Drop-in ``Adapter`` (mirror of reference src/modeling/models/adapter.py) whose forward/backward
run as hand-written sm_100a kernels behind the C ABI.

Kept from the reference: constructor signature, the ``adapter_{i}_down`` /
``adapter_{i}_up`` ``nn.Linear`` sub-modules (same state-dict keys and shapes, so main.py's
substring selection of 'adapter_0' / 'adapter_1' / 'adapter_2' keeps working), ``set_active_adapter``
with its ``requires_grad`` toggling (adapter.py:66-95), ``activate_gating`` / ``deactivate_gating``,
``forward(hidden_states, input_tensor)``, ``adapter_layer_forward_bert``, attributes ``gating``,
``scaling``, ``actv``.  New, optional: ``rank=`` (the reference parses --adapter_reduction_factor but
never forwards it) and ``activation=`` ('relu' = reference, adapter.py:24).

There is no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import ops
from .._lib import FeddatError


def init_bert_weights(module):
    """adapter.py:5-14: N(0, 0.02) weights, zero biases, unit LayerNorm."""
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=0.02)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


_PENDING = object()      # marks a parameter gradient that the deferred weight-gradient queue will deliver


class _DatFunction(torch.autograd.Function):
    """y = res + scale * (act(x Wd^T + bd) Wu^T + bu) over the concatenated active branches.

    inputs: x [M, 768] bf16, res (same tensor object as x when the residual IS the input), then the
    fp32 master parameters (down_w, down_b, up_w, up_b) of every active branch.
    """

    @staticmethod
    def forward(ctx, adapter: "Adapter", names, x, res, *params):
        need_bwd = any(ctx.needs_input_grad[2:])
        segs = adapter._segments(params, need_bwd=need_bwd, key=names)
        same = res is x
        y = None
        # ReLU: keep the hidden (what autograd would save) so that backward skips the recompute of
        # x Wd^T; GELU's derivative needs the pre-activation, which the backward kernel recomputes
        save_h = need_bwd and adapter._act_code == ops.ACT_RELU
        hs = []
        for i, (pk, _, _) in enumerate(segs):
            out = ops.dat_forward(x, res if i == 0 else y, pk, adapter._scale(), adapter._act_code,
                                  save_hidden=save_h)
            y, h = out if save_h else (out, None)
            hs.append(h)
        ctx.hs = hs
        ctx.adapter = adapter
        ctx.scale = adapter._scale()        # mode at FORWARD time (backward may run after a mode switch)
        ctx.segs = segs
        ctx.same = same
        ctx.n_params = len(params)
        ctx.param_needs = [p.requires_grad for p in params]
        ctx.params = params            # leaves: the deferred weight-gradient queue assigns their .grad itself
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        adapter = ctx.adapter
        dy = dy.contiguous()
        need_dx = ctx.needs_input_grad[2]
        need_dres = ctx.needs_input_grad[3] and not ctx.same
        r = adapter.rank
        nb = ctx.n_params // 4
        # which branches train (adapter.py:71-85 toggles requires_grad per mode)
        branch_trains = [any(ctx.param_needs[4 * b: 4 * b + 4]) for b in range(nb)]
        grads: List[Optional[torch.Tensor]] = [None] * ctx.n_params
        dx_total = None
        # Deferred weight gradients (ops.DeferredWgrad: the trainer computes all sites' gradients in a few launches
        # after the backward pass) when every trained branch lies whole inside one segment, so that its gradient
        # tensors are plain views of one launch's outputs (true for every width the train step uses: a branch is
        # only cut when r > 256)
        queue = ops.deferred_queue()
        defer = queue is not None and all(
            not branch_trains[b] or any(c0 <= b * r and (b + 1) * r <= c0 + w for _, c0, w in ctx.segs) for b in range(nb))
        for si, (pk, col0, width) in enumerate(ctx.segs):
            # trainable slice of this segment in concatenated-bottleneck coordinates
            lo, hi = None, None
            for b in range(nb):
                b0, b1 = max(b * r, col0), min((b + 1) * r, col0 + width)
                if branch_trains[b] and b0 < b1:
                    lo = b0 if lo is None else min(lo, b0)
                    hi = b1 if hi is None else max(hi, b1)
            ts = None if lo is None else (lo - col0, hi - col0)
            dx, g = ops.dat_backward(x, dy, pk, ctx.scale, adapter._act_code, train_slice=ts,
                                     need_dx=need_dx, add_dy=(ctx.same and si == 0), hidden=ctx.hs[si],
                                     allow_defer=defer)
            if dx is not None:
                dx_total = dx if dx_total is None else dx_total + dx
            if g is not None:
                d_down_w, d_down_b, d_up_w, d_up_b = g
                for b in range(nb):
                    b0, b1 = max(b * r, lo), min((b + 1) * r, hi)
                    if not branch_trains[b] or b0 >= b1:
                        continue
                    s0, s1 = b0 - lo, b1 - lo          # rows/cols inside the slice gradients
                    p0, p1 = b0 - b * r, b1 - b * r    # rows/cols inside the branch parameters
                    full = (p0 == 0 and p1 == r)

                    def put(idx, piece, shape, sl):
                        if defer:                 # filled at the flush, handed to .grad there
                            if ctx.param_needs[idx]:
                                queue.assign(ctx.params[idx], piece)
                            grads[idx] = _PENDING
                        elif full:
                            grads[idx] = piece.contiguous() if grads[idx] is None else grads[idx] + piece
                        else:
                            if grads[idx] is None:
                                grads[idx] = torch.zeros(shape, device=x.device, dtype=torch.float32)
                            grads[idx][sl] += piece

                    put(4 * b + 0, d_down_w[s0:s1], (r, adapter.model_dim), slice(p0, p1))
                    put(4 * b + 1, d_down_b[s0:s1], (r,), slice(p0, p1))
                    put(4 * b + 2, d_up_w[:, s0:s1], (adapter.model_dim, r), (slice(None), slice(p0, p1)))
                    if si == 0 or grads[4 * b + 3] is None:
                        # d_up_b is the same column sum of dY for every segment: take it once
                        if grads[4 * b + 3] is None:
                            if defer:
                                put(4 * b + 3, d_up_b, None, None)
                            else:
                                grads[4 * b + 3] = d_up_b.contiguous()
        for i, needs in enumerate(ctx.param_needs):
            if not needs or grads[i] is _PENDING:
                grads[i] = None
        dres = dy if need_dres else None
        return (None, None, dx_total if need_dx else None, dres, *grads)


class _DualDatFunction(torch.autograd.Function):
    """Two DAT modes over the two halves of ONE row-stacked tensor (``Adapter.set_dual``): rows [0, M/2) see
    the gating pair (adapter_0 | adapter_2, scale 0.5), rows [M/2, M) see adapter_1 alone -- the MKD
    schedule's passes A/C and B sharing every frozen-backbone launch.  ONE grouped kernel launch per direction
    (forward; backward data gradient; backward weight gradients) covers both halves, each half with its own
    packed weights, bottleneck width and scale, writing into the halves of one output tensor (no
    concatenation copies); residual == input.

    inputs: x [M, 768] bf16, then the 8 gating parameters (adapter_0, adapter_2) and the 4 of adapter_1.
    """

    @staticmethod
    def forward(ctx, adapter: "Adapter", x, *params):
        need_bwd = any(ctx.needs_input_grad[1:])
        half = x.shape[0] // 2
        specs = ((params[:8], 0.5 * adapter.scaling, ("adapter_0", "adapter_2")), (params[8:], 1.0, ("adapter_1",)))
        save_h = need_bwd and adapter._act_code == ops.ACT_RELU
        y = torch.empty_like(x)
        groups, meta = [], []
        for i, (ps, scale, key) in enumerate(specs):
            (pk, _, _), = adapter._segments(ps, need_bwd=need_bwd, key=key)
            xs = x[i * half:(i + 1) * half]
            groups.append(dict(x=xs, res=xs, w=pk, scale=scale, out=y[i * half:(i + 1) * half], save_hidden=save_h))
            meta.append((pk, scale, [p.requires_grad for p in ps]))
        outs = ops.dat_forward_grouped(groups, adapter._act_code)
        ctx.parts = [(pk, scale, h, needs) for (pk, scale, needs), (_, h) in zip(meta, outs)]
        ctx.adapter = adapter
        ctx.params = params            # leaves: the deferred weight-gradient queue assigns their .grad itself
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        adapter = ctx.adapter
        dy = dy.contiguous()
        need_dx = ctx.needs_input_grad[1]
        half = x.shape[0] // 2
        r = adapter.rank
        dx = torch.empty_like(dy) if need_dx else None
        groups, slices = [], []
        for i, (pk, scale, hid, needs) in enumerate(ctx.parts):
            sl = slice(i * half, (i + 1) * half)
            nb = len(needs) // 4
            trains = [any(needs[4 * b: 4 * b + 4]) for b in range(nb)]
            lo = next((b * r for b in range(nb) if trains[b]), None)
            hi = None if lo is None else max((b + 1) * r for b in range(nb) if trains[b])
            slices.append((lo, hi, trains, nb))
            groups.append(dict(x=x[sl], dy=dy[sl], w=pk, scale=scale, train_slice=None if lo is None else (lo, hi),
                               need_dx=need_dx, add_dy=True, hidden=hid, dx_out=None if dx is None else dx[sl]))
        results = ops.dat_backward_grouped(groups, adapter._act_code, allow_defer=True)
        queue = ops.deferred_queue()     # set: the gradient tensors are filled when the trainer flushes the queue
        grads = []
        for (_, g), (lo, hi, trains, nb), (_, _, _, needs) in zip(results, slices, ctx.parts):
            part = [None] * len(needs)
            if g is not None:
                d_down_w, d_down_b, d_up_w, d_up_b = g
                for b in range(nb):
                    if trains[b]:
                        s0, s1 = b * r - lo, (b + 1) * r - lo
                        part[4 * b:4 * b + 4] = [d_down_w[s0:s1], d_down_b[s0:s1], d_up_w[:, s0:s1], d_up_b]
            if queue is not None:
                first = len(grads)
                for k, (gr, nd) in enumerate(zip(part, needs)):
                    if nd and gr is not None:
                        queue.assign(ctx.params[first + k], gr)
                grads += [None] * len(needs)
            else:
                grads += [gr.contiguous() if (nd and gr is not None) else None for gr, nd in zip(part, needs)]
        return (None, dx, *grads)


class Adapter(nn.Module):
    """The DAT bottleneck operator (reference adapter.py:16-163)."""

    def __init__(self, names, device, model_dim=768, adapter_reduction_factor=16, rank=None,
                 activation="relu"):
        super().__init__()
        if activation not in ("relu", "gelu"):
            raise ValueError(f"activation must be 'relu' (reference) or 'gelu', got {activation!r}")
        self.actv = nn.ReLU() if activation == "relu" else nn.GELU()
        self._act_code = ops.act_code(activation)
        self.scaling = 1.0
        self.gating = False
        self.model_dim = model_dim
        self.rank = int(rank) if rank is not None else model_dim // adapter_reduction_factor
        if model_dim != 768:
            raise FeddatError(f"the sm_100a DAT kernels are built for model_dim=768 (got {model_dim})")
        if self.rank % 16 != 0 or self.rank < 16:
            raise FeddatError(f"adapter rank must be a multiple of 16 (got {self.rank}); "
                              "there is no fallback path for other ranks")

        if isinstance(names, str):
            names = [names]
        self.adapter_dict = {}
        for name in names:
            if "adapter" in name:
                for part, (fan_in, fan_out) in (("down", (model_dim, self.rank)), ("up", (self.rank, model_dim))):
                    n = f"{name}_{part}"
                    setattr(self, n, nn.Linear(fan_in, fan_out).to(device))
                    m = getattr(self, n)
                    m.apply(init_bert_weights)
                    for p in m.parameters():
                        p.requires_grad = True
            elif name in ["gating"]:
                # kept for state-dict compatibility; the learned gate is commented out upstream
                # (adapter.py:143) and never used in forward
                setattr(self, f"{name}_module", nn.Linear(model_dim, 2).to(device))
                m = getattr(self, f"{name}_module")
                m.apply(init_bert_weights)

        if hasattr(self, "adapter_2_down"):                      # adapter.py:55-58
            for m in [self.adapter_2_down, self.adapter_2_up]:
                for p in m.parameters():
                    p.requires_grad = False
        self._active_name: Optional[str] = None
        self._dual = False
        self._bank = {}                 # branch-name tuple -> segments, packed by refresh_packs()
        self._bank_store = {}           # persistent operand buffers behind the bank
        self._bank_valid = False

    # ------------------------------------------------------------------ mode switches (reference API)
    def deactivate_gating(self):
        self.gating = False

    def activate_gating(self):
        self.gating = True

    def set_active_adapter(self, name):
        """adapter.py:66-95, including the requires_grad side effects."""
        if isinstance(name, str):
            self.active_adapter_down = getattr(self, f"{name}_down")
            self.active_adapter_up = getattr(self, f"{name}_up")
            self._active_name = name

        if name == "adapter_0":
            self._set_grad(("adapter_0",), True)
            self._set_grad(("adapter_1",), False)
        elif name == "adapter_1":
            self._set_grad(("adapter_1",), True)
            self._set_grad(("adapter_0",), False)
        elif isinstance(name, list):
            self._set_grad(name, True)
        return

    def _set_grad(self, names: Sequence[str], flag: bool):
        for n in names:
            for part in ("down", "up"):
                m = getattr(self, f"{n}_{part}", None)
                if m is not None:
                    for p in m.parameters():
                        p.requires_grad = flag

    def set_dual(self, on: bool):
        """Row-batched MKD mode (TaskTrainer's batched schedule): the first half of the rows goes through the
        gating pair, the second half through adapter_1 (see _DualDatFunction).  Both adapter_0 and
        adapter_1 are trainable in this mode; adapter_2 stays frozen."""
        if on:
            if not hasattr(self, "adapter_2_down"):
                raise FeddatError("dual mode needs the three DAT adapters (adapter_0, adapter_1, adapter_2)")
            if 2 * self.rank > ops.MAX_R_TOTAL:
                raise FeddatError("dual mode covers ranks up to 128 (gating width 256)")
            self._set_grad(("adapter_0", "adapter_1"), True)
        self._dual = bool(on)

    # ------------------------------------------------------------------ kernel plumbing
    def _scale(self) -> float:
        return 0.5 * self.scaling if self.gating else 1.0      # adapter.py:144-146 / :131

    def _active_branch_names(self) -> Tuple[str, ...]:
        if not self.gating:
            if self._active_name is None:
                raise AttributeError("set_active_adapter() was never called")  # reference: adapter.py:127
            return (self._active_name,)
        if hasattr(self, "adapter_2_down"):
            return ("adapter_0", "adapter_2")                   # adapter.py:135
        return ("adapter_0", "adapter_1")                       # adapter.py:151

    def _branch_params(self, names):
        out = []
        for n in names:
            d, u = getattr(self, f"{n}_down"), getattr(self, f"{n}_up")
            out += [d.weight, d.bias, u.weight, u.bias]
        return out

    def _seg_specs(self, params, need_bwd: bool = True):
        """Packing jobs of the given branches' parameters, split into <= 256-wide launches:
        [(ops.PackSpec, first column, width)].  Segments of a wide bottleneck are row / column SLICES of the
        masters (no copies); the (summed) up bias rides on the first segment only."""
        for p in params:
            if not (p.is_cuda and p.dtype == torch.float32):
                raise FeddatError("adapter parameters must be fp32 CUDA tensors (fp32 masters; the "
                                  "kernels consume a packed bf16 copy)")
        nb = len(params) // 4
        r = self.rank
        br = [[params[4 * b + i].detach() for i in range(4)] for b in range(nb)]
        if nb * r <= ops.MAX_R_TOTAL:
            return [(ops.PackSpec([b[0] for b in br], [b[1] for b in br], [b[2] for b in br],
                                  [b[3] for b in br], need_bwd), 0, nb * r)]
        out, col = [], 0
        for b in range(nb):
            dw, db, uw, _ = br[b]
            for j0 in range(0, r, ops.MAX_R_TOTAL):
                j1 = min(r, j0 + ops.MAX_R_TOTAL)
                bu_src = [x[3] for x in br] if col == 0 else [None, None]
                out.append((ops.PackSpec([dw[j0:j1]], [db[j0:j1]], [uw[:, j0:j1]], bu_src, need_bwd), col, j1 - j0))
                col += j1 - j0
        return out

    def _segments(self, params, need_bwd: bool = True, key=None):
        """Packed bf16 operands of the active branches: [(PackedWeights, first column, width)].

        Inside a train step the trainer has packed every site once (``refresh_packs`` below: ONE launch for
        all sites and both modes) and the operands are taken from that bank.  Outside -- eval, ad-hoc calls --
        they are packed afresh on every forward: nothing observable from Python tells reliably that a
        parameter's values changed (``torch._fused_adamw_`` does not bump ``Tensor._version`` and neither
        does the ``state_dict()[k].data.copy_(...)`` idiom of the reference round loop, main.py:446-450,
        task_trainer.py:36-41), so a cache without an owner that knows when weights change would go stale
        silently.  Backward reuses the forward's packs through ``ctx.segs``."""
        if self._bank_valid and key is not None and key in self._bank:
            return self._bank[key]
        specs = self._seg_specs(params, need_bwd)
        with torch.no_grad():
            packs = ops.pack_weights_batched([sp for sp, _, _ in specs])
        return [(pk, c0, w) for pk, (_, c0, w) in zip(packs, specs)]

    def bank_keys(self):
        """The branch sets a train step uses: the gating pair and adapter_1 alone (task_trainer.py:280-330)."""
        if not hasattr(self, "adapter_1_down"):
            return []
        gate = ("adapter_0", "adapter_2") if hasattr(self, "adapter_2_down") else ("adapter_0", "adapter_1")
        return [gate, ("adapter_1",)]

    # ------------------------------------------------------------------ forward (reference API)
    def forward(self, hidden_states, input_tensor):
        """adapter.py:124-163."""
        if not hidden_states.is_cuda:
            raise FeddatError("Adapter.forward: CPU tensors are not supported -- the DAT operator exists "
                              "only as sm_100a CUDA kernels (no fallback)")
        shape = hidden_states.shape
        out_dtype = hidden_states.dtype
        if self._dual:
            if input_tensor is not hidden_states:
                raise FeddatError("dual mode is defined for sites whose residual is the input (ViLT, ViT)")
            x = hidden_states.reshape(-1, shape[-1])
            if x.dtype != torch.bfloat16:
                x = x.to(torch.bfloat16)
            x = x.contiguous()
            if x.shape[0] % 2:
                raise FeddatError("dual mode needs an even number of rows (two stacked copies of the batch)")
            params = self._branch_params(("adapter_0", "adapter_2", "adapter_1"))
            y = _DualDatFunction.apply(self, x, *params).view(shape)
            return y if out_dtype == torch.bfloat16 else y.to(out_dtype)
        names = self._active_branch_names()
        params = self._branch_params(names)
        same = input_tensor is hidden_states
        x = hidden_states.reshape(-1, shape[-1])
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        x = x.contiguous()
        if same:
            res = x
        else:
            res = input_tensor.reshape(-1, shape[-1])
            if res.dtype != torch.bfloat16:
                res = res.to(torch.bfloat16)
            res = res.contiguous()
        y = _DatFunction.apply(self, names, x, res, *params)
        y = y.view(shape)
        return y if out_dtype == torch.bfloat16 else y.to(out_dtype)

    # ------------------------------------------------------------------ BERT-site wrapper (reference API)
    def adapter_layer_forward_bert(self, hidden_states, input_tensor, layer_norm):
        hidden_states, residual = self.pre_forward(hidden_states, input_tensor, layer_norm)
        hidden_states = self.forward(hidden_states, residual)
        hidden_states = self.post_forward(hidden_states, input_tensor, layer_norm)
        return hidden_states

    def pre_forward(self, hidden_states, input_tensor, layer_norm):
        residual = hidden_states                                # adapter.py:104 residual_before_ln
        if layer_norm:
            from .fused_ln import layer_norm_of_sum             # add + LayerNorm in one launch where it applies
            hidden_states = layer_norm_of_sum(layer_norm, hidden_states, input_tensor)
        else:
            hidden_states = hidden_states + input_tensor
        return hidden_states, residual

    def post_forward(self, hidden_states, input_tensor, layer_norm):
        if layer_norm:
            from .fused_ln import layer_norm_of_sum
            hidden_states = layer_norm_of_sum(layer_norm, hidden_states, input_tensor)
        else:
            hidden_states = hidden_states + input_tensor
        return hidden_states


def refresh_packs(adapters: Sequence[Adapter]) -> None:
    """Packs the operands of every given site for both train-step modes in ONE launch
    (feddat_pack_weights_batched) into persistent buffers and marks the sites' banks valid.  The trainer calls
    this at the start of a train step -- every forward of the MKD schedule runs before the optimizer step that
    changes the branches it reads (task_trainer.py:280-330: pass B updates adapter_1 and the head, pass C
    adapter_0 and the head; adapter_2 is frozen) -- and ``invalidate_packs`` at its end."""
    specs, outs, slots = [], [], []
    with torch.no_grad():
        for a in adapters:
            for key in a.bank_keys():
                segs = a._seg_specs(a._branch_params(key), need_bwd=True)
                store = a._bank_store.get(key)
                if store is None:
                    dev = segs[0][0].down_w[0].device
                    store = a._bank_store[key] = [ops.alloc_packed(sp.down_w[0].shape[0], len(sp.down_w), a.model_dim, dev)
                                                  for sp, _, _ in segs]
                for (sp, c0, w), pk in zip(segs, store):
                    specs.append(sp); outs.append(pk)
                slots.append((a, key, [(pk, c0, w) for (_, c0, w), pk in zip(segs, store)]))
        ops.pack_weights_batched(specs, outs)
    for a, key, segs in slots:
        a._bank[key] = segs
        a._bank_valid = True


def invalidate_packs(adapters: Sequence[Adapter]) -> None:
    for a in adapters:
        a._bank_valid = False

"""Fused (residual add +) LayerNorm for the frozen ViLT blocks around each DAT site, as sm_100a kernels
behind the C ABI (``feddat_ln_fwd`` / ``feddat_ln_bwd``, csrc/layernorm.cu).

``fast_vilt_layer_forward`` restates HF ``ViltLayer.forward`` (the caller of the reference's
``Adaptered_ViltOutput.forward``, src/modeling/adaptered_output.py:73-79) with the two LayerNorms and the
first residual connection fused:

    ln1        = LN_before(h)                         one launch
    attn       = attention(ln1)                        (HF module, SDPA)
    h2, ln2    = h + attn, LN_after(h + attn)          one launch for the add AND the LayerNorm (it also
                                                       emits h2 + b_out, the output dense layer's bias)
    inter      = GELU(dense(ln2))                      cuBLAS + feddat_gelu_*
    x          = (h2 + b_out) + inter W_out^T          ONE GEMM with a beta = 1 epilogue (the second residual)
    out        = adapter(x, x)                         the DAT site

It applies only where the kernels are defined -- bf16 CUDA activations of width 768, frozen bf16 affine
parameters (main.py:138-139 freezes the whole backbone), no attention mask / attention maps requested --
and otherwise calls the stock HF forward, so nothing changes for other configurations.
"""
from __future__ import annotations

import logging
import types

import torch

from .. import ops

_warned = set()
FUSE_MLP_GELU = True      # A/B switch (tests, bench): False = cuBLAS GEMM + streaming GELU kernels


def _fallback_once(what: str, why: str) -> None:
    """A frozen-backbone fast path that silently falls through to stock HF code would hide a performance
    regression: say so once per cause."""
    if (what, why) not in _warned:
        _warned.add((what, why))
        logging.getLogger("feddat_b200").warning("%s: stock HF path taken (%s)", what, why)


class _AddLayerNorm(torch.autograd.Function):
    """y = LayerNorm(s), s = x + res (``res`` may be None: s is x).  Returns y alone (no res), (y, s), or --
    with ``bias2`` -- (y, s + bias2): the residual stream with the next dense layer's bias pre-added, the
    only form in which it is then consumed.  Affine parameters and bias2 are frozen: no gradients for them."""

    @staticmethod
    def forward(ctx, x, res, weight, bias, eps, bias2):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        r2 = None if res is None else res.reshape(-1, shape[-1])
        y, s, mean, rstd, s2 = ops.layer_norm_fwd(x2, r2, weight, bias, eps, bias2)
        ctx.save_for_backward(s, weight, mean, rstd)
        ctx.has_res = res is not None
        ctx.shape = shape
        if bias2 is not None:
            return y.view(shape), s2.view(shape)
        if res is None:
            return y.view(shape)
        return y.view(shape), s.view(shape)

    @staticmethod
    def backward(ctx, gy, gs=None):
        s, weight, mean, rstd = ctx.saved_tensors
        d = s.shape[-1]
        gy2 = gy.reshape(-1, d).contiguous()
        gs2 = None if gs is None else gs.reshape(-1, d).contiguous()
        dx = ops.layer_norm_bwd(gy2, gs2, s, weight, mean, rstd).view(ctx.shape)
        return dx, (dx if ctx.has_res else None), None, None, None, None


class _LayerNormPass(torch.autograd.Function):
    """(LayerNorm(x), x) for frozen affine parameters.  The second output IS x: the caller routes the residual stream
    through it, so both gradients reaching x -- through the LayerNorm and through the residual add further down -- meet
    in this node's backward and leave as ONE launch (feddat_ln_bwd with ``dsum``) instead of a LayerNorm backward plus
    autograd's gradient-sum kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        shape = x.shape
        y, _, mean, rstd, _ = ops.layer_norm_fwd(x.reshape(-1, shape[-1]), None, weight, bias, eps, None)
        ctx.save_for_backward(x, weight, mean, rstd)
        return y.view(shape), x.view_as(x)

    @staticmethod
    def backward(ctx, gy, gx):
        x, weight, mean, rstd = ctx.saved_tensors
        d = x.shape[-1]
        gx2 = None if gx is None else gx.reshape(-1, d).contiguous()
        dx = ops.layer_norm_bwd(gy.reshape(-1, d).contiguous(), gx2, x.reshape(-1, d), weight, mean, rstd)
        return dx.view(x.shape), None, None, None


class _Gelu(torch.autograd.Function):
    """Exact-erf GELU (HF ViltIntermediate's activation) through feddat_gelu_fwd / feddat_gelu_bwd."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.gelu_fwd(x)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return ops.gelu_bwd(gy.contiguous(), x)


class _FrozenMlp(torch.autograd.Function):
    """h = res_b + gelu(ln2 W_in^T + b_in) W_out^T for the FROZEN MLP of a ViLT block (HF ViltIntermediate +
    ViltOutput.dense; res_b already carries the second dense layer's bias).  The exact GELU rides in the epilogue of
    this repo's tcgen05 GEMM (feddat_mlp_fc1_gelu_fwd) and its derivative in the epilogue of the backward's first
    GEMM (feddat_mlp_fc2_dgelu_bwd); the two 768-wide products stay cuBLAS.  No weight gradients: everything here
    is frozen (main.py:138-139)."""

    @staticmethod
    def forward(ctx, ln2, res_b, w_in, b_in, w_out, w_out_t):
        shape = res_b.shape
        pre, act = ops.mlp_fc1_gelu(ln2.reshape(-1, ln2.shape[-1]), w_in, b_in)
        h = torch.addmm(res_b.reshape(-1, shape[-1]), act, w_out.t())
        ctx.save_for_backward(pre, w_in, w_out_t)
        ctx.shape = shape
        return h.view(shape)

    @staticmethod
    def backward(ctx, dh):
        pre, w_in, w_out_t = ctx.saved_tensors
        dh2 = dh.reshape(-1, dh.shape[-1]).contiguous()
        dpre = ops.mlp_fc2_dgelu(dh2, w_out_t, pre)
        d_ln2 = torch.mm(dpre, w_in)
        return d_ln2.view(*ctx.shape[:-1], w_in.shape[1]), dh, None, None, None, None


def _transposed_weight(dense: torch.nn.Linear) -> torch.Tensor:
    """[in, out] copy of a FROZEN Linear's [out, in] weight (K-major B operand of the backward GEMM), cached on the
    module and rebuilt when the weight's storage or version changes (load_state_dict, dtype casts)."""
    w = dense.weight
    key = (w.data_ptr(), w._version, w.dtype)
    cached = getattr(dense, "_feddat_wt", None)
    if cached is None or cached[0] != key:
        with torch.no_grad():
            cached = (key, w.detach().t().contiguous())
        dense._feddat_wt = cached
    return cached[1]


def _bias_f32(dense: torch.nn.Linear) -> torch.Tensor:
    """fp32 copy of a FROZEN Linear's bias (what the fused epilogue adds to its fp32 accumulator), cached like
    ``_transposed_weight``."""
    b = dense.bias
    key = (b.data_ptr(), b._version, b.dtype)
    cached = getattr(dense, "_feddat_b32", None)
    if cached is None or cached[0] != key:
        with torch.no_grad():
            cached = (key, b.detach().float().contiguous())
        dense._feddat_b32 = cached
    return cached[1]


def _mlp_ok(inter_mod, dense: torch.nn.Linear, h: torch.Tensor) -> bool:
    """The fused GEMM + GELU path: exact-erf GELU, frozen bf16 Linear layers with biases, tile-aligned widths."""
    act = inter_mod.intermediate_act_fn
    exact_gelu = type(act).__name__ == "GELUActivation" or (isinstance(act, torch.nn.GELU) and act.approximate == "none")
    d1 = getattr(inter_mod, "dense", None)
    return (exact_gelu and isinstance(d1, torch.nn.Linear) and d1.bias is not None and d1.weight.dtype == torch.bfloat16
            and h.dtype == torch.bfloat16 and h.is_cuda and not d1.weight.requires_grad and not d1.bias.requires_grad
            and d1.weight.is_contiguous() and dense.weight.is_contiguous()
            and d1.out_features % 256 == 0 and d1.in_features % 64 == 0 and dense.in_features == d1.out_features)


def intermediate(mod, h: torch.Tensor) -> torch.Tensor:
    """HF ViltIntermediate.forward: dense (cuBLAS) + GELU (this repo's streaming kernel when the module's
    activation is the exact GELU and the tensor is CUDA bf16)."""
    act = mod.intermediate_act_fn
    exact_gelu = type(act).__name__ == "GELUActivation" or (isinstance(act, torch.nn.GELU) and act.approximate == "none")
    if not exact_gelu:
        return mod(h)
    pre = mod.dense(h)
    if pre.is_cuda and pre.dtype == torch.bfloat16 and pre.is_contiguous() and pre.numel() % 8 == 0:
        return _Gelu.apply(pre)
    return act(pre)


def _usable(ln: torch.nn.LayerNorm, h: torch.Tensor) -> bool:
    return (h.is_cuda and h.dtype == torch.bfloat16 and h.shape[-1] == 768 and h.is_contiguous()
            and tuple(ln.normalized_shape) == (768,) and ln.weight is not None and ln.bias is not None
            and ln.weight.dtype == torch.bfloat16 and not ln.weight.requires_grad and not ln.bias.requires_grad)


def layer_norm(ln: torch.nn.LayerNorm, h: torch.Tensor) -> torch.Tensor:
    if _usable(ln, h):
        return _AddLayerNorm.apply(h, None, ln.weight, ln.bias, ln.eps, None)
    return ln(h)


def layer_norm_pass(ln: torch.nn.LayerNorm, h: torch.Tensor):
    """(LayerNorm(h), h): take the residual stream from the second value (see ``_LayerNormPass``)."""
    if _usable(ln, h) and h.requires_grad and torch.is_grad_enabled():
        return _LayerNormPass.apply(h, ln.weight, ln.bias, ln.eps)
    return layer_norm(ln, h), h


def add_layer_norm(ln: torch.nn.LayerNorm, a: torch.Tensor, b: torch.Tensor, bias2=None):
    """(a + b [+ bias2], LayerNorm(a + b))."""
    if _usable(ln, a) and b.dtype == a.dtype and b.shape == a.shape and b.is_contiguous():
        y, s = _AddLayerNorm.apply(a, b, ln.weight, ln.bias, ln.eps, bias2)
        return s, y
    s = a + b
    return (s if bias2 is None else s + bias2), ln(s)


def layer_norm_of_sum(ln, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """LayerNorm(a + b) -- the post-LN form of BERT's output blocks (xbert.py:350-361, 428-445) and of the DAT wrapper
    around them (adapter.py:97-116) -- as ONE launch where the kernels apply, ``ln(a + b)`` otherwise."""
    if (isinstance(ln, torch.nn.LayerNorm) and _usable(ln, a) and b.dtype == a.dtype and b.shape == a.shape
            and b.is_contiguous()):
        return _AddLayerNorm.apply(a, b, ln.weight, ln.bias, ln.eps, None)[0]
    return ln(a + b)


def _prebias_ok(out_mod, h: torch.Tensor) -> bool:
    """The "dense + bias + residual" of Adaptered_ViltOutput can be ONE GEMM (beta = 1 epilogue) when its
    dense layer is frozen bf16 with a bias and its dropout is the identity."""
    dense = getattr(getattr(out_mod, "layer", None), "dense", None)
    drop = getattr(getattr(out_mod, "layer", None), "dropout", None)
    return (hasattr(out_mod, "adapter") and isinstance(dense, torch.nn.Linear) and dense.bias is not None
            and dense.weight.dtype == h.dtype and not dense.weight.requires_grad and not dense.bias.requires_grad
            and dense.bias.is_contiguous() and (drop is None or drop.p == 0.0 or not out_mod.training))


def fast_vilt_layer_forward(self, hidden_states, attention_mask=None, output_attentions=False):
    """HF ViltLayer.forward with fused LayerNorms (see module docstring)."""
    if attention_mask is not None or output_attentions or not _usable(self.layernorm_before, hidden_states):
        _fallback_once("fast_vilt_layer_forward",
                       "attention mask given" if attention_mask is not None else
                       "attention maps requested" if output_attentions else
                       f"activations {hidden_states.dtype} on {hidden_states.device.type} or unfrozen / non-bf16 LayerNorm")
        return type(self).forward(self, hidden_states, attention_mask, output_attentions)
    ln1, hidden_states = layer_norm_pass(self.layernorm_before, hidden_states)
    attention_output = self.attention(ln1, None, output_attentions=False)[0]
    if _prebias_ok(self.output, hidden_states):
        # first residual + LayerNorm in one launch, which also emits the residual stream with the output
        # dense layer's bias pre-added; second residual = the GEMM's beta = 1 epilogue; then the DAT site
        dense = self.output.layer.dense
        res_b, ln2 = add_layer_norm(self.layernorm_after, attention_output, hidden_states, bias2=dense.bias)
        if FUSE_MLP_GELU and _mlp_ok(self.intermediate, dense, ln2) and ln2.is_contiguous():
            # first GEMM + exact GELU (and, in backward, the second GEMM's data gradient + GELU') in one kernel each
            d1 = self.intermediate.dense
            h = _FrozenMlp.apply(ln2, res_b, d1.weight, _bias_f32(d1), dense.weight, _transposed_weight(dense))
            return (self.output.adapter(h, h),)
        _fallback_once("fast_vilt_layer_forward", "MLP not a frozen bf16 768 -> 3072 -> 768 pair with the exact GELU: "
                                                   "GEMM + streaming GELU kernels")
        inter = intermediate(self.intermediate, ln2)
        h = torch.addmm(res_b.reshape(-1, res_b.shape[-1]), inter.reshape(-1, inter.shape[-1]), dense.weight.t())
        h = h.view(res_b.shape)
        return (self.output.adapter(h, h),)
    _fallback_once("fast_vilt_layer_forward", "output dense layer not a frozen bf16 Linear with bias / active dropout: "
                                               "second residual stays a separate add")
    hidden_states, ln2 = add_layer_norm(self.layernorm_after, attention_output, hidden_states)   # first residual
    layer_output = intermediate(self.intermediate, ln2)
    layer_output = self.output(layer_output, hidden_states)                                     # second residual + DAT
    return (layer_output,)


def enable(vilt_model) -> None:
    """Patch every encoder layer of an HF ViltModel (instance-level, like ViltEncoderWrapper.enable_sdpa)."""
    for layer in vilt_model.encoder.layer:
        layer.forward = types.MethodType(fast_vilt_layer_forward, layer)

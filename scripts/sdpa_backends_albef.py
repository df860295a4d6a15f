"""Which torch SDPA backend is fastest for ALBEF's BERT-side attention shapes (bf16, additive masks, dropout 0.1)?
Device time of 20 calls queued behind a spin kernel (the ALBEF step is replayed from a graph: host time does not count)."""
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

dev = "cuda"
SHAPES = {"text self  (16,12,25,25)": (16, 25, 25, "pad"), "text cross (16,12,25,577)": (16, 25, 577, None),
          "dec self   (32,12,6,6)": (32, 6, 6, "causal"), "dec cross  (32,12,6,25)": (32, 6, 25, "pad"),
          "vit        (16,12,577,577)": (16, 577, 577, "vit")}


def graph_time(fn, n=20):
    """device time per call: the calls are queued behind a ~10 ms spin kernel, so the GPU runs them back to back"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(20_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for label, (b, sq, sk, mk) in SHAPES.items():
    q0 = torch.randn(b, sq, 12, 64, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k0 = torch.randn(b, sk, 12, 64, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v0 = torch.randn(b, sk, 12, 64, device=dev, dtype=torch.bfloat16, requires_grad=True)
    q, k, v = q0.transpose(1, 2), k0.transpose(1, 2), v0.transpose(1, 2)
    go = torch.randn(b, 12, sq, 64, device=dev, dtype=torch.bfloat16)
    mask, p = None, 0.1
    if mk == "pad":
        mask = torch.zeros(b, 1, 1, sk, device=dev, dtype=torch.bfloat16)
        mask[:, :, :, sk - 3:] = -10000.0
    elif mk == "causal":
        mask = torch.triu(torch.full((sq, sk), -10000.0, device=dev, dtype=torch.bfloat16), 1)[None, None].expand(b, 1, sq, sk).contiguous()
    elif mk == "vit":
        p = 0.0
    for name, be in (("default", None), ("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION),
                     ("efficient", SDPBackend.EFFICIENT_ATTENTION), ("math", SDPBackend.MATH)):
        try:
            def fb():
                o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=p)
                o.backward(go)
                q0.grad = k0.grad = v0.grad = None

            def f():
                with torch.no_grad():
                    F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=p)
            if be is None:
                tf, tb = graph_time(f), graph_time(fb)
            else:
                with sdpa_kernel(be):
                    tf, tb = graph_time(f), graph_time(fb)
            print(f"{label:28s} {name:10s} fwd {tf:7.1f} us   fwd+bwd {tb:7.1f} us", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{label:28s} {name:10s} unavailable: {type(e).__name__}: {str(e)[:70]}", flush=True)

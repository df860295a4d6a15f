"""Which torch SDPA backend is fastest for the frozen ViLT attention shape (B=32, H=12, S=185, d=64, bf16)?"""
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

dev = "cuda"
q, k, v = (torch.randn(32, 12, 185, 64, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
go = torch.randn(32, 12, 185, 64, device=dev, dtype=torch.bfloat16)


def bench(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION),
                 ("efficient", SDPBackend.EFFICIENT_ATTENTION), ("math", SDPBackend.MATH)):
    try:
        with sdpa_kernel(be):
            def fwd():
                with torch.no_grad():
                    return F.scaled_dot_product_attention(q, k, v)

            def fwdbwd():
                o = F.scaled_dot_product_attention(q, k, v)
                o.backward(go)
            tf, tb = bench(fwd), bench(fwdbwd)
        print(f"{name:10s} fwd {tf:7.1f} us   fwd+bwd {tb:7.1f} us")
    except Exception as e:  # noqa: BLE001
        print(f"{name:10s} unavailable: {type(e).__name__}: {str(e)[:80]}")
with torch.no_grad():
    pass
# default dispatch
print("default    fwd %.1f us" % bench(lambda: F.scaled_dot_product_attention(q.detach(), k.detach(), v.detach())))

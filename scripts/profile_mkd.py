"""The fused ALBEF MKD head alone at the bench's logits shape, for ncu:
    ncu --set full --clock-control none -k regex:mkd -o gpurun_out/r2_mkd python scripts/profile_mkd.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
n_seq, La, C = 33, 6, 30522
lab = torch.randint(1000, C, (n_seq, La), device=dev, generator=g)
lab[:, 4:] = -100
w = torch.rand(n_seq, device=dev, generator=g) / 16
for dt in (torch.bfloat16, torch.float32):
    for _ in range(3):
        sc = torch.randn(n_seq, La, C, device=dev, generator=g).to(dt)
        te = torch.randn(n_seq, La, C, device=dev, generator=g).to(dt)
        ops.mkd_ce_loss(sc, te[:, :-1], lab, w, 2.0)
torch.cuda.synchronize()
print("ok")

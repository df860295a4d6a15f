"""Timeline of CTA 0's pipeline events inside dat_fused_kernel (debug aid).
    FEDDAT_DEBUG_LIB=1 python scripts/trace_kernel.py [R] [M] [fwd|bwd|bwdw|gfwd|gbwd]
(gfwd / gbwd: the grouped launch over [M gating rows (R) | M adapter_1 rows (R / 2)] of one site; the trace hooks
exist only in the -DFEDDAT_DEBUG twin of the library, hence the environment switch)"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import _lib, ops  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = int(sys.argv[2]) if len(sys.argv) > 2 else 71040
mode = sys.argv[3] if len(sys.argv) > 3 else "fwd"
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
nb = 2 if R > 128 else 1
r = R // nb
pk = ops.pack_weights([[torch.randn(r, 768, device=dev, generator=g) * 0.02, torch.zeros(r, device=dev),
                        torch.randn(768, r, device=dev, generator=g) * 0.02, torch.zeros(768, device=dev)]
                       for _ in range(nb)])
x = torch.randn(M, 768, device=dev, generator=g).to(torch.bfloat16)
dy = torch.randn(M, 768, device=dev, generator=g).to(torch.bfloat16)
lib = _lib.load_debug()
pk1 = ops.pack_weights([[torch.randn(r, 768, device=dev, generator=g) * 0.02, torch.zeros(r, device=dev),
                         torch.randn(768, r, device=dev, generator=g) * 0.02, torch.zeros(768, device=dev)]])
x1 = torch.randn(M, 768, device=dev, generator=g).to(torch.bfloat16)
if mode.startswith("g"):
    (_, h2), (_, h1) = ops.dat_forward_grouped([dict(x=x, res=x, w=pk, scale=0.5, save_hidden=True),
                                                dict(x=x1, res=x1, w=pk1, scale=1.0, save_hidden=True)])


def run():
    if mode == "gfwd":
        ops.dat_forward_grouped([dict(x=x, res=x, w=pk, scale=0.5, save_hidden=True),
                                 dict(x=x1, res=x1, w=pk1, scale=1.0, save_hidden=True)])
    elif mode == "gbwd":
        ops.dat_backward_grouped([dict(x=x, dy=dy, w=pk, scale=0.5, train_slice=(0, r), hidden=h2),
                                  dict(x=x1, dy=dy, w=pk1, scale=1.0, train_slice=(0, r), hidden=h1)])
    elif mode == "fwd":
        ops.dat_forward(x, x, pk, 0.5)
    else:
        ops.dat_backward(x, dy, pk, 0.5, train_slice=(0, r) if mode == "bwdw" else None, need_dx=True)


for _ in range(3):
    run()
buf = torch.zeros(1024, dtype=torch.int64, device=dev)      # 256 events of CTA 0 + entry / exit stamps of every CTA
lib.feddat_debug_set_trace(_lib.ptr(buf))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
lib.feddat_debug_set_trace(None)
t = buf.cpu().tolist()
t0 = t[1] if t[1] else t[0]
names = {1: "kernel entry", 2: "CTA exit", 0: "start", 22: "G1 issued", 23: "G1b issued", 24: "h_full passed", 40: "epiA: P full", 41: "epiA: E1 done",
         110: "prod: tile start", 111: "prod: G1 slots issued", 112: "prod: all slots issued"}
for k in range(12):
    names[10 + k] = f"mma: G1 kc{k} slots ready"
for c in range(6):
    names[25 + c] = f"mma: chunk{c} D buffer free"
    names[31 + c] = f"mma: chunk{c} issued"
    names[42 + 4 * c] = f"epi{c & 1}: chunk{c} D full"
    names[43 + 4 * c] = f"epi{c & 1}: chunk{c}.0 res ready"
    names[44 + 4 * c] = f"epi{c & 1}: chunk{c}.1 res ready"
    names[45 + 4 * c] = f"epi{c & 1}: chunk{c} done"
    names[104 + c] = f"store: chunk{c}.x issued"
for k in range(12):
    names[90 + k] = f"res: load c64={k} issued"
for c in range(2):
    for w in range(8):
        names[70 + 8 * c + w] = f"E1: CTA {c} warp {4 + w} (group {'AB'[w >> 2]}) hidden packed"
names.update({120: "E2 c0.0: D full seen", 121: "E2 c0.0: tmem_ld issued", 122: "E2 c0.0: residual ready",
              123: "E2 c0.0: tmem_ld done", 124: "E2 c0.0: math + st.shared done", 125: "E2 c0.0: proxy fence done",
              126: "E2 c0.0: arrived"})
wg = {200: "wgrad: entry", 201: "wgrad: prologue done", 202: "wgrad: all MMAs issued", 203: "wgrad: accumulators complete",
      204: "wgrad: split barrier passed (all partials stored)", 206: "wgrad: final slices written", 205: "wgrad: exit"}
for i in range(10):
    wg[210 + i] = f"wgrad: stage {i} loads issued"
    wg[220 + i] = f"wgrad: stage {i} MMAs issued"
    wg[230 + i] = f"wgrad: stage {i} bias sums done (warp 2)"
print(f"kernel {e0.elapsed_time(e1) * 1e3:.1f} us, R={R} M={M} {mode}")
if any(t[e] for e in wg):
    w0 = t[200]
    for e in sorted((e for e in wg if t[e]), key=lambda e: t[e]):
        print(f"wgrad {(t[e] - w0) / 1e3:9.2f} us  {wg[e]}   (dgrad entry {-(t[1] - w0) / 1e3:.2f} us earlier)" if e == 200 else
              f"wgrad {(t[e] - w0) / 1e3:9.2f} us  {wg[e]}")
for tile in range(2):
    ev = [(t[tile * 128 + e] - t0, e) for e in range(128) if t[tile * 128 + e] != 0]
    for dt, e in sorted(ev):
        print(f"tile{tile} {dt / 1e3:9.2f} us  {names.get(e, e)}")

ctas = [(t[256 + 2 * b], t[257 + 2 * b]) for b in range(384) if t[256 + 2 * b]]
if ctas:
    first = min(a for a, _ in ctas)
    ent = sorted((a - first) / 1e3 for a, _ in ctas)
    ext = sorted((b - first) / 1e3 for _, b in ctas)
    dur = sorted((b - a) / 1e3 for a, b in ctas)
    q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]      # noqa: E731
    print(f"{len(ctas)} CTAs: entry min/median/max {ent[0]:.2f} / {q(ent, .5):.2f} / {ent[-1]:.2f} us after the first; "
          f"exit min/median/max {ext[0]:.2f} / {q(ext, .5):.2f} / {ext[-1]:.2f} us; "
          f"CTA time min/median/max {dur[0]:.2f} / {q(dur, .5):.2f} / {dur[-1]:.2f} us")
    slow = sorted(range(len(ctas)), key=lambda i: -(ctas[i][1] - first))[:6]
    print("last to exit:", [(i, round((ctas[i][0] - first) / 1e3, 2), round((ctas[i][1] - first) / 1e3, 2)) for i in slow])

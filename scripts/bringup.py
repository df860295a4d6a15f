"""GPU bring-up: runs each tcgen05 probe / kernel case in its own subprocess (a trap in one case must
not poison the CUDA context of the next) and prints one JSON line per case.

    python scripts/bringup.py            # all cases
    python scripts/bringup.py --case probe:0:0:256:128
"""
from __future__ import annotations

import argparse
import ctypes
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def run_probe(a_mode, b_mode, N, K, overrides=None):
    import torch
    from feddat_b200 import _lib
    lib = _lib.load_debug()
    torch.manual_seed(0)
    A = torch.randn(128, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    A_in = A.t().contiguous() if a_mode == 1 else A.contiguous()
    B_in = B.t().contiguous() if b_mode == 1 else B.contiguous()
    D = torch.full((128, N), float("nan"), device="cuda", dtype=torch.float32)
    ov = None
    if overrides:
        ov = (ctypes.c_uint32 * 6)(*overrides)
    rc = lib.feddat_probe_gemm(_lib.ptr(A_in), _lib.ptr(B_in), _lib.ptr(D), N, K, a_mode, b_mode, ov,
                               _lib.stream_ptr())
    _lib.check(rc, "probe")
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    return {"max_abs_err": err, "ref_absmax": ref.abs().max().item(), "ok": bool(err < 1e-2)}


def run_pair(N, K, a_tmem, reps=1):
    import torch
    from feddat_b200 import _lib
    lib = _lib.load_debug()
    torch.manual_seed(0)
    A = torch.randn(256, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    D = torch.full((256, N), float("nan"), device="cuda", dtype=torch.float32)
    ns = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(lib.feddat_probe_pair(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), N, K, a_tmem, reps, _lib.ptr(ns),
                                     _lib.stream_ptr()))
    torch.cuda.synchronize()
    if reps > 1:
        n_mma = reps * K // 16
        t = ns.item()
        return {"n_mma": n_mma, "ns_per_mma": t / n_mma, "tflops_pair": 2 * 256 * N * 16 * n_mma / t / 1e3,
                "ok": True}
    err = (D - ref).abs()
    return {"max_abs_err": err.max().item(), "err_top": err[:128].max().item(), "err_bot": err[128:].max().item(),
            "nan": int(torch.isnan(D).sum().item()), "ok": bool(err.max().item() < 1e-2)}


def run_fwd(R, M, scale, act=0, alias=True):
    import torch
    from feddat_b200 import _lib
    lib = _lib.load_debug()
    torch.manual_seed(1)
    d = 768
    X = torch.randn(M, d, device="cuda").to(torch.bfloat16)
    Res = X if alias else torch.randn(M, d, device="cuda").to(torch.bfloat16)
    Wd = (torch.randn(R, d, device="cuda") * 0.05).to(torch.bfloat16)
    Wu = (torch.randn(d, R, device="cuda") * 0.05).to(torch.bfloat16)
    bd = torch.randn(R, device="cuda") * 0.1
    bu = torch.randn(d, device="cuda") * 0.1
    Y = torch.full((M, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.feddat_dat_fwd(_lib.ptr(X), _lib.ptr(Res), _lib.ptr(Y), _lib.ptr(Wd), _lib.ptr(bd),
                            _lib.ptr(Wu), _lib.ptr(bu), M, d, R, scale, act, 0, _lib.stream_ptr())
    _lib.check(rc, "dat_fwd")
    torch.cuda.synchronize()
    P = X.float() @ Wd.float().t() + bd
    H = torch.relu(P) if act == 0 else torch.nn.functional.gelu(P)
    H = H.to(torch.bfloat16).float()
    ref = Res.float() + scale * (H @ Wu.float().t() + bu)
    err = (Y.float() - ref).abs().max().item()
    nan = int(torch.isnan(Y.float()).sum().item())
    # timing
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        lib.feddat_dat_fwd(_lib.ptr(X), _lib.ptr(Res), _lib.ptr(Y), _lib.ptr(Wd), _lib.ptr(bd),
                           _lib.ptr(Wu), _lib.ptr(bu), M, d, R, scale, act, 0, _lib.stream_ptr())
    ev0.record()
    n = 20
    for _ in range(n):
        lib.feddat_dat_fwd(_lib.ptr(X), _lib.ptr(Res), _lib.ptr(Y), _lib.ptr(Wd), _lib.ptr(bd),
                           _lib.ptr(Wu), _lib.ptr(bu), M, d, R, scale, act, 0, _lib.stream_ptr())
    ev1.record()
    torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1e3 / n
    return {"max_abs_err": err, "nan": nan, "ok": bool(err < 6e-2 and nan == 0), "us": us,
            "GBs": 4 * d * M / us / 1e3, "TFs": 4 * d * R * M / us / 1e6}


CASES = [
    "probe:0:0:64:64", "probe:0:0:256:128", "probe:0:0:48:128",
    "probe:0:1:256:128", "probe:1:0:256:128", "probe:1:1:256:128", "probe:1:1:128:64",
    "probe:2:0:256:128", "probe:2:0:64:64",
    "fwd:48:128:1.0", "fwd:128:300:0.5", "fwd:256:5920:0.5", "fwd:96:5920:0.5", "fwd:256:71040:0.5",
    "fwd:96:71040:0.5", "fwd:48:71040:1.0", "fwd:128:71040:0.5", "fwd:64:1000:1.0:1",
]


def run_case(case: str):
    parts = case.split(":")
    if parts[0] == "probe":
        a, b, N, K = map(int, parts[1:5])
        ov = list(map(int, parts[5:11])) if len(parts) >= 11 else None
        return run_probe(a, b, N, K, ov)
    if parts[0] == "pair":
        return run_pair(int(parts[1]), int(parts[2]), int(parts[3]), int(parts[4]) if len(parts) > 4 else 1)
    if parts[0] == "fwd":
        R, M = int(parts[1]), int(parts[2])
        scale = float(parts[3])
        act = int(parts[4]) if len(parts) > 4 else 0
        return run_fwd(R, M, scale, act)
    if parts[0] == "bwd":
        from scripts import bringup_bwd
        return bringup_bwd.run_case(parts)
    raise ValueError(case)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--cases", nargs="*", default=None)
    ap.add_argument("--timeout", type=int, default=120)
    args = ap.parse_args()
    if args.case:
        try:
            res = run_case(args.case)
        except Exception as e:  # noqa: BLE001
            res = {"ok": False, "error": f"{type(e).__name__}: {e}"}
        print("RESULT " + json.dumps({"case": args.case, **res}), flush=True)
        return
    for case in (args.cases or CASES):
        try:
            r = subprocess.run([sys.executable, __file__, "--case", case], capture_output=True,
                               text=True, timeout=args.timeout)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                print(lines[-1][7:], flush=True)
            else:
                print(json.dumps({"case": case, "ok": False, "rc": r.returncode,
                                  "stdout": r.stdout[-600:], "stderr": r.stderr[-1200:]}), flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": case, "ok": False, "error": "timeout"}), flush=True)


if __name__ == "__main__":
    main()

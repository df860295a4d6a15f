"""Benchmark of the FedDAT per-client train step (BASELINE.json: "VQA samples/sec/box (ViLT+DAT
bf16)").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

A "step" is one ``TaskTrainer.train_step`` (3 forwards, 2 backwards, 2 AdamW+scheduler steps:
reference task_trainer.py:280-330) over one synthetic batch of BASELINE config 1: ViLT-B/32 + DAT
rank 128, 384x384 image / 40 text tokens, batch 32, MKD temperature 2.0, bf16 backbone.  With N GPUs
every rank trains its own client on its own batch (weak scaling) and the timed region ends with the
round-boundary FedAvg allreduce of the flat adapter_1 buffer.

One JSON line is printed by rank 0 (see DESIGN.md "Measurement" for every key).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "VQA samples/sec/box (ViLT+DAT bf16)"
WORKLOAD = "ViLT-B32 + DAT rank-128 bf16, synthetic VQA 384x384 / 40-tok batch=32, MKD tau=2.0 (BASELINE configs[1])"
B, T, H, C, RANK, TEMP = 32, 40, 384, 100, 128, 2.0
D = 768


def read_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"hbm_gbs": j["hbm_gbs"], "tf_burst": j["bf16_tflops"], "tf_sustained": j["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.nvml, self.handle, self.samples, self._stop = None, None, [], False

    # NVML in a thread (a sample every ~4 ms: a 20-step timed region is only 0.12 s, the nvidia-smi loop below delivers
    # one line per 100 ms); the nvidia-smi loop stays as the fallback when pynvml is not importable
    def _start_nvml(self) -> bool:
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    idx = self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:  # noqa: BLE001
            self.nvml = None
            return False
        threading.Thread(target=self._pump_nvml, daemon=True).start()
        return True

    def _pump_nvml(self):
        n = self.nvml
        while not self._stop:
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    why = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:  # noqa: BLE001
                    why = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((time.time(), mhz, why))
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.004)

    def start(self):
        if self._start_nvml():
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def _stop_nvml(self, t0, t1):
        n = self.nvml
        self._stop = True
        rows = [r for r in self.samples if t0 <= r[0] <= t1] or self.samples
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted(name for name, bit in bits.items() if any(r[2] & bit for r in rows))
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            mx = None
        sm = [float(r[1]) for r in rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(sm), "source": "nvml, one sample per ~4 ms inside the timed region"}

    def stop(self, t0, t1):
        if self.nvml is not None:
            return self._stop_nvml(t0, t1)
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------
def build_client(rank: int, device):
    """Model + trainer for one federated client on this rank's GPU (random-init ViLT-B/32, seeded)."""
    import torch
    import torch.nn as nn
    from feddat_b200.modeling.vilt import convert_batch_to_vilt_input_dict
    from feddat_b200.train.accelerator import Accelerator
    from feddat_b200.train.fedavg import FlatCommBuffer
    from feddat_b200.train.prepare import default_args, place_on_gpu, prepare_model
    from feddat_b200.train.task_trainer import TaskTrainer, get_polynomial_decay_schedule_with_warmup

    torch.manual_seed(1000 * 1)                          # identical backbone / global adapter on every rank
    task = f"synth{rank % 8}"
    args = default_args(ordered_cl_tasks=[task], adapter_rank=RANK)
    model = prepare_model(args, place=False)
    place_on_gpu(model, device)
    comm = FlatCommBuffer(model, model.comm_state_dict_names)
    tr = TaskTrainer()
    tr.args = SimpleNamespace(optimizer_mode="dat", encoder_name="vilt", debug=0)
    tr.accelerator = Accelerator(device=device)
    tr.device = torch.device(device)
    tr.task_key = task
    tr.batch2inputs_converter = convert_batch_to_vilt_input_dict
    tr.loss_criterion = nn.BCEWithLogitsLoss(reduction="mean")
    tr.weight_decay, tr.lr, tr.adam_epsilon, tr.kl_temp = 1e-2, 1e-4, 1e-8, TEMP
    sd = model.state_dict()
    for name in sd:                                      # task_trainer.py:36-45
        if "adapter_1" in name:
            sd[name.replace("adapter_1", "adapter_2")].data.copy_(sd[name].data)
    for n, p in model.named_parameters():
        if "adapter_2" in n:
            p.requires_grad = False
    wrapped = tr.accelerator.prepare(model)
    opt = tr.create_optimizer(wrapped)
    sched = get_polynomial_decay_schedule_with_warmup(opt, 100, 1000, lr_end=0, power=1)
    wrapped.train()
    return SimpleNamespace(model=model, wrapped=wrapped, trainer=tr, opt=opt, sched=sched, comm=comm, task=task)


def kernel_rooflines(peaks, device):
    """Per-kernel CUDA-event timing of the DAT kernels at the step's shapes (M = 32 x 185 rows per
    site) and at the 12-site batched size (steady state).  Cold inputs without a write-flush: every
    launch reads the next of a rotation of input sets totalling > 2x the 126 MB L2 (a 256 MB memset
    between launches leaves the L2 full of dirty lines whose write-back then competes with the
    kernel's own HBM reads).  A spin kernel ahead of each timed launch absorbs the host-side launch
    latency (tensor-map encodes), so the events bracket GPU time only."""
    import torch
    from feddat_b200 import ops
    out = {}
    g = torch.Generator(device=device).manual_seed(0)

    def mk(r, nb):
        return ops.pack_weights([[torch.randn(r, D, device=device, generator=g) * 0.02,
                                  torch.zeros(r, device=device),
                                  torch.randn(D, r, device=device, generator=g) * 0.02,
                                  torch.zeros(D, device=device)] for _ in range(nb)])

    def timeit(fn, sets, iters=12):
        for i in range(3):
            fn(*sets[i % len(sets)])
        ts = []
        for i in range(iters):
            args_i = sets[(3 + i) % len(sets)]
            torch.cuda._sleep(1_000_000)                    # ~0.5 ms: every launch of fn is queued behind it (the
                                                            # backward path spends ~80 us of host time in allocations, tensor-map encodes and ctypes calls)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*args_i)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return statistics.mean(ts)

    for M in (B * 185, 12 * B * 185):
        n_sets = max(2, -(-300_000_000 // (2 * M * D * 2)))
        if M == B * 185:
            n_sets = max(n_sets, 24)                        # 12 sites x 2 row groups for the deferred weight gradients
        pk2, pk1 = mk(RANK, 2), mk(RANK, 1)
        r = RANK
        sets = []
        for _ in range(n_sets):
            x = torch.randn(M, D, device=device, generator=g).to(torch.bfloat16)
            dy = torch.randn(M, D, device=device, generator=g).to(torch.bfloat16)
            # the hidden the forward saves for the backward (ReLU path: no recompute in dgrad)
            h2 = ops.dat_forward(x, x, pk2, 0.5, save_hidden=True)[1]
            h1 = ops.dat_forward(x, x, pk1, 1.0, save_hidden=True)[1]
            sets.append((x, dy, h2, h1))
        cases = {
            "fwd_gating": (lambda x, dy, h2, h1: ops.dat_forward(x, x, pk2, 0.5, save_hidden=True), 8 * D * r * M, 4 * D * M),
            "fwd_single": (lambda x, dy, h2, h1: ops.dat_forward(x, x, pk1, 1.0, save_hidden=True), 4 * D * r * M, 4 * D * M),
            "bwd_gating": (lambda x, dy, h2, h1: ops.dat_backward(x, dy, pk2, 0.5, train_slice=(0, r), hidden=h2),
                           12 * D * r * M, 6 * D * M),
            "bwd_single": (lambda x, dy, h2, h1: ops.dat_backward(x, dy, pk1, 1.0, train_slice=(0, r), hidden=h1),
                           8 * D * r * M, 6 * D * M),
        }
        for name, (fn, flops, nbytes) in cases.items():
            t = timeit(fn, sets)
            t_tensor, t_hbm = flops / (peaks["tf_burst"] * 1e12), nbytes / (peaks["hbm_gbs"] * 1e9)
            bound = "tensor" if t_tensor >= t_hbm else "hbm"
            out[f"{name}_M{M}"] = {
                "us": round(t * 1e6, 2), "bound": bound, "tflops": round(flops / t / 1e12, 1),
                "gbs": round(nbytes / t / 1e9, 1), "frac_of_roofline": round(max(t_tensor, t_hbm) / t, 4)}
        if M == B * 185:
            # what the train step actually launches per adapter site (batched MKD schedule): ONE grouped launch
            # over [gating rows | adapter_1 rows] per direction
            Mh = M
            gsets = []
            for (x, dy, h2, h1), (x1, dy1, _, h1b) in zip(sets[0::2], sets[1::2]):
                gsets.append((x, dy, h2, x1, dy1, h1b, torch.empty_like(x), torch.empty_like(x1)))

            def g_fwd(x, dy, h2, x1, dy1, h1, y0, y1):
                ops.dat_forward_grouped([dict(x=x, res=x, w=pk2, scale=0.5, out=y0, save_hidden=True),
                                         dict(x=x1, res=x1, w=pk1, scale=1.0, out=y1, save_hidden=True)])

            def g_bwd(x, dy, h2, x1, dy1, h1, y0, y1):
                ops.dat_backward_grouped([dict(x=x, dy=dy, w=pk2, scale=0.5, train_slice=(0, r), hidden=h2, dx_out=y0),
                                          dict(x=x1, dy=dy1, w=pk1, scale=1.0, train_slice=(0, r), hidden=h1, dx_out=y1)])

            def g_dgrad(*a):
                # what a site's backward launches inside the train step: the data gradient only -- the weight
                # gradients of all sites are ONE launch after the backward pass (ops.DeferredWgrad)
                with ops.deferred_wgrad() as q:
                    ops.dat_backward_grouped([dict(x=a[0], dy=a[1], w=pk2, scale=0.5, train_slice=(0, r), hidden=a[2], dx_out=a[6]),
                                              dict(x=a[3], dy=a[4], w=pk1, scale=1.0, train_slice=(0, r), hidden=a[5], dx_out=a[7])],
                                             allow_defer=True)
                    q.groups, q.keep = [], []

            for name, fn, flops, nbytes in (("site_fwd_grouped", g_fwd, 12 * D * r * Mh, 4 * D * 2 * Mh),
                                            ("site_bwd_grouped", g_bwd, 20 * D * r * Mh, 6 * D * 2 * Mh),
                                            ("site_dgrad_grouped", g_dgrad, 12 * D * r * Mh, 4 * D * 2 * Mh)):
                t = timeit(fn, gsets)
                t_tensor, t_hbm = flops / (peaks["tf_burst"] * 1e12), nbytes / (peaks["hbm_gbs"] * 1e9)
                out[f"{name}_M{2 * Mh}"] = {
                    "us": round(t * 1e6, 2), "bound": "tensor" if t_tensor >= t_hbm else "hbm",
                    "tflops": round(flops / t / 1e12, 1), "gbs": round(nbytes / t / 1e9, 1),
                    "frac_of_roofline": round(max(t_tensor, t_hbm) / t, 4),
                    "rows": f"{Mh} gating (R = {2 * r}) + {Mh} adapter_1 (R = {r})"}
            # the deferred weight-gradient launch over 12 sites (24 groups x 6 column chunks = 144 CTAs, no row
            # splits): 12 data-gradient launches queue their groups, the flush is timed alone (0.5 GB of X / dY / H /
            # dP: cold by construction)
            n_sites = min(12, len(gsets))
            ts = []
            for it in range(6):
                with ops.deferred_wgrad() as q:
                    for a in gsets[:n_sites]:
                        ops.dat_backward_grouped([dict(x=a[0], dy=a[1], w=pk2, scale=0.5, train_slice=(0, r), hidden=a[2], dx_out=a[6]),
                                                  dict(x=a[3], dy=a[4], w=pk1, scale=1.0, train_slice=(0, r), hidden=a[5], dx_out=a[7])],
                                                 allow_defer=True)
                    torch.cuda._sleep(1_000_000)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    q.flush()
                    e1.record()
                    torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1) * 1e-3)
            t = statistics.mean(ts)
            wbytes = n_sites * 2 * Mh * (2 * D + 2 * r) * 2         # X, dY, H_t, dP_t of both row groups
            wflops = n_sites * 2 * Mh * 4 * D * r
            out[f"wgrad_deferred_{n_sites}sites"] = {
                "us": round(t * 1e6, 2), "us_per_site": round(t * 1e6 / n_sites, 2), "bound": "hbm",
                "tflops": round(wflops / t / 1e12, 1), "gbs": round(wbytes / t / 1e9, 1),
                "frac_of_roofline": round(wbytes / t / 1e9 / peaks["hbm_gbs"], 4),
                "bytes": "X, dY (768 columns) + H_t, dP_t (128 columns) of both row groups of every site, read once",
                "grid": f"{2 * n_sites} groups x 6 column chunks, one CTA each over all {Mh} rows"}
            del gsets
        del sets
    return out


def backbone_kernel_timings(device, batch):
    """The two own kernels that replaced library ops in the frozen block around each site, next to what they replaced
    (same shapes as the step: 2 x batch sequences of 185 tokens; CUDA events, rotating input sets):
    GEMM + exact GELU (feddat_mlp_fc1_gelu_fwd / feddat_mlp_fc2_dgelu_bwd) vs cuBLAS + this repo's streaming GELU, and the
    short-sequence attention (feddat_attn_fwd / feddat_attn_bwd) vs F.scaled_dot_product_attention (cuDNN flash)."""
    import torch
    import torch.nn.functional as F
    from feddat_b200 import ops
    g = torch.Generator(device=device).manual_seed(2)
    Bq, S, H, Dh = 2 * batch, 185, 12, 64
    M = Bq * S
    w1 = (torch.randn(4 * D, D, device=device, generator=g) * 0.05).to(torch.bfloat16)
    b1 = (torch.randn(4 * D, device=device, generator=g) * 0.1).to(torch.bfloat16)
    w2t = (torch.randn(4 * D, D, device=device, generator=g) * 0.05).to(torch.bfloat16)
    sets = []
    for _ in range(5):
        a, dy, q, k, v, do = (torch.randn(M, D, device=device, generator=g).to(torch.bfloat16) for _ in range(6))
        sets.append((a, dy, torch.addmm(b1, a, w1.t()), q.view(Bq, S, H, Dh), k.view(Bq, S, H, Dh), v.view(Bq, S, H, Dh),
                     do.view(Bq, S, H, Dh)))
    b1f = b1.float()

    def timeit(fn, iters=10):
        for i in range(3):
            fn(*sets[i % 5])
        ts = []
        for i in range(iters):
            torch.cuda._sleep(1_000_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(*sets[(3 + i) % 5]); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return round(sum(ts) / len(ts), 2)

    o_l = [ops.attn_fwd(s_[3], s_[4], s_[5], 0.125) for s_ in sets]

    def sdpa_fb(a, dy, pre, q, k, v, do):
        q, k, v = (t.detach().permute(0, 2, 1, 3).requires_grad_(True) for t in (q, k, v))
        F.scaled_dot_product_attention(q, k, v).backward(do.permute(0, 2, 1, 3))

    own = {
        "mlp_fc1_gelu_fwd": timeit(lambda a, dy, pre, q, k, v, do: ops.mlp_fc1_gelu(a, w1, b1f)),
        "mlp_fc2_dgelu_bwd": timeit(lambda a, dy, pre, q, k, v, do: ops.mlp_fc2_dgelu(dy, w2t, pre)),
        "attn_fwd": timeit(lambda a, dy, pre, q, k, v, do: ops.attn_fwd(q, k, v, 0.125)),
        "attn_bwd": timeit(lambda a, dy, pre, q, k, v, do: ops.attn_bwd(do, q, k, v, o_l[0][0], o_l[0][1], 0.125)),
    }
    sdpa_f = timeit(lambda a, dy, pre, q, k, v, do: F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3),
                                                                                 v.permute(0, 2, 1, 3)))
    lib = {
        "mlp_fc1_gelu_fwd": timeit(lambda a, dy, pre, q, k, v, do: ops.gelu_fwd(torch.addmm(b1, a, w1.t()))),
        "mlp_fc2_dgelu_bwd": timeit(lambda a, dy, pre, q, k, v, do: ops.gelu_bwd(torch.mm(dy, w2t.t()), pre)),
        "attn_fwd": sdpa_f,
        "attn_bwd": round(timeit(sdpa_fb) - sdpa_f, 2),
    }
    what = {"mlp_fc1_gelu_fwd": "cuBLAS addmm + streaming GELU", "mlp_fc2_dgelu_bwd": "cuBLAS mm + streaming GELU'",
            "attn_fwd": "F.scaled_dot_product_attention", "attn_bwd": "its autograd backward (fwd + bwd minus fwd)"}
    shape = {"mlp_fc1_gelu_fwd": f"[{M}, {D}] x [{D}, {4 * D}]", "mlp_fc2_dgelu_bwd": f"[{M}, {D}] x [{D}, {4 * D}]",
             "attn_fwd": f"B={Bq} S={S} H={H} d={Dh}", "attn_bwd": f"B={Bq} S={S} H={H} d={Dh}"}
    return {k_: {"us": own[k_], "library_us": lib[k_], "vs_library": round(lib[k_] / own[k_], 2), "library": what[k_],
                 "shape": shape[k_]} for k_ in own}


def eager_reference_adapter(peaks, device):
    """The reference operator itself on this GPU ("the bar is PyTorch-eager (cuBLAS) on the
    same box"): the arithmetic of reference adapter.py:124-163 as PyTorch eager ops -- fp32 nn.Linear masters
    under bf16 autocast (what accelerate's mixed precision does to it), the (B, S, 2) tensor of 0.5s built per
    call, the per-branch unsqueeze-multiply aggregation -- forward and autograd backward, at the same shapes
    and with the same cold-input rotation as ``kernel_rooflines``.  adapter_0 trains, adapter_2 is frozen
    (gating); adapter_1 trains (single)."""
    import torch
    import torch.nn as nn
    out = {}
    g = torch.Generator(device=device).manual_seed(1)

    def lin(i, o):
        m = nn.Linear(i, o).to(device)
        with torch.no_grad():
            m.weight.normal_(0, 0.02, generator=g); m.bias.zero_()
        return m

    ad = {n: (lin(D, RANK), lin(RANK, D)) for n in ("adapter_0", "adapter_1", "adapter_2")}
    for m in ad["adapter_2"]:
        for p in m.parameters():
            p.requires_grad = False
    relu = nn.ReLU()

    def fwd(x, gating):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if not gating:                                          # adapter.py:125-131
                down, up = ad["adapter_1"]
                return x + up(relu(down(x)))
            ups = [ad[n][1](relu(ad[n][0](x))) for n in ("adapter_0", "adapter_2")]     # :135-141
            w = (torch.ones(x.shape[0], x.shape[1], 2) * 0.5).to(device)                # :144 (host-built constant)
            agg = w[:, :, 0].unsqueeze(-1) * ups[0] + w[:, :, 1].unsqueeze(-1) * ups[1]  # :118-122
            return x + agg * 1.0                                    # :146

    for M in (B * 185, 12 * B * 185):
        n_sets = max(2, -(-300_000_000 // (2 * M * D * 2)))
        sets = [(torch.randn(M // 185, 185, D, device=device, generator=g).to(torch.bfloat16),
                 torch.randn(M // 185, 185, D, device=device, generator=g).to(torch.bfloat16)) for _ in range(n_sets)]
        for gating in (True, False):
            tf, tb = [], []
            for i in range(3 + 10):
                x, dy = sets[i % n_sets]
                x = x.clone().requires_grad_(True)
                torch.cuda._sleep(1_000_000)
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                y = fwd(x, gating)
                e[1].record()
                y.backward(dy)
                e[2].record()
                torch.cuda.synchronize()
                for m in ad.values():
                    for mm in m:
                        mm.zero_grad(set_to_none=True)
                if i >= 3:
                    tf.append(e[0].elapsed_time(e[1]) * 1e3); tb.append(e[1].elapsed_time(e[2]) * 1e3)
            key = "gating" if gating else "single"
            out[f"fwd_{key}_M{M}"] = round(statistics.mean(tf), 2)
            out[f"bwd_{key}_M{M}"] = round(statistics.mean(tb), 2)
        del sets
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from feddat_b200 import ops
    from feddat_b200.synthetic import make_vilt_batch, to_device
    from feddat_b200.train.fedavg import get_average_net_flat

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)"
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = read_peaks()
    c = build_client(rank, device)
    K, W = args.steps, args.warmup

    # K + W distinct host batches in pinned memory (seed = 1000 * config + rank)
    host = [make_vilt_batch(B, T, H, C, seed=(1000 * 1 + rank) * 1000 + i, client=rank % 8, pin=True)
            for i in range(min(K + W, 8))]
    dev_batches = [to_device(b, device) for b in host]
    torch.cuda.synchronize()

    if args.eager:
        def run_step(batch):
            return c.trainer.train_step(c.wrapped, 0, batch, c.opt, c.sched)
    else:
        # the public graphed step (feddat_b200/train/graphed.py): two eager steps, then one capture of
        # the whole train_step (3 fwd / 2 bwd / 2 AdamW), replayed per batch
        from feddat_b200.train.graphed import GraphedTrainStep
        run_step = GraphedTrainStep(c.trainer, c.wrapped, c.opt, c.sched, dev_batches[0], warmup=2)

    def step_resident(i):
        return run_step(dev_batches[i % len(dev_batches)])

    def step_e2e(i):
        # every step: H2D of this step's inputs from pinned host memory + D2H read of its result
        if args.eager:
            batch = to_device(host[i % len(host)], device)
            return run_step(batch).item()
        # graphed: batch i was staged by prefetch() while step i-1 computed; stage batch i+1 now, so the
        # copy engine works under this step's compute (all K copies still happen inside the timed region)
        loss = run_step()
        run_step.prefetch(host[(i + 1) % len(host)])
        return loss.item()

    def round_boundary():
        if world > 1:
            get_average_net_flat(c.comm, [c.comm.flat], [1.0], total=float(world))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(n):
            fn(i)
        round_boundary()
        e1.record()
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, t0, t1

    for i in range(W):
        step_resident(i)
    # the round-boundary collective is warmed like every other kernel (the first ncclAllReduce of a
    # communicator sets up its channels: 22 ms in round 1's N = 8 run) and timed on its own
    allreduce_us = None
    if world > 1:
        for _ in range(2):
            round_boundary()
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(5):
            round_boundary()
        eb.record()
        torch.cuda.synchronize()
        t = torch.tensor([ea.elapsed_time(eb) / 5 * 1e3], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allreduce_us = round(t.item(), 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = ops.launch_count
    ms, t0, t1 = timed(step_resident, K)
    launches = ops.launch_count - l0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    if not args.eager:
        run_step.prefetch(host[0])
    for i in range(min(W, 2)):
        step_e2e(i)
    ms_e2e, _, _ = timed(step_e2e, K)

    enc0 = host[0]["encodings"]
    copied = list(enc0.keys()) if args.eager else run_step._used_keys(enc0)
    h2d = sum(enc0[k].numel() * enc0[k].element_size() for k in copied if hasattr(enc0[k], "numel")) + \
        host[0]["target_scores"].numel() * host[0]["target_scores"].element_size()
    result = None
    if rank == 0:
        value = world * B * K / (ms * 1e-3)
        e2e = world * B * K / (ms_e2e * 1e-3)
        kr = kernel_rooflines(peaks, device)
        eager = eager_reference_adapter(peaks, device)
        try:
            kr.update(backbone_kernel_timings(device, B))
        except Exception as exc:                              # the headline numbers do not depend on this leg
            kr["backbone_kernels_error"] = repr(exc)[:200]
        for name, us in eager.items():
            if name in kr:
                kr[name]["eager_reference_us"] = us
                kr[name]["vs_eager"] = round(us / kr[name]["us"], 2)
        for d_ in ("fwd", "bwd"):                       # the grouped site launch vs the two eager reference calls
            k_ = f"site_{d_}_grouped_M{2 * B * 185}"
            if k_ in kr:
                us = eager[f"{d_}_gating_M{B * 185}"] + eager[f"{d_}_single_M{B * 185}"]
                kr[k_]["eager_reference_us"] = round(us, 2)
                kr[k_]["vs_eager"] = round(us / kr[k_]["us"], 2)
        # Dominant DAT work of the step = the backward of one adapter site in the batched MKD schedule, as the step
        # launches it: ONE grouped data-gradient launch per site over [5920 gating rows (R = 256) | 5920 adapter_1
        # rows (R = 128)] + that site's 1/12 share of the ONE deferred weight-gradient launch over all 12 sites.
        # Algorithmic work: 12 d r + 8 d r = 20 d r FLOP per row pair, 6 d bytes per row
        # (read X, read dY, write dX) -> HBM-bound at the measured peaks.
        Mh = B * 185
        pair = kr[f"site_bwd_grouped_M{2 * Mh}"]            # dgrad + per-site wgrad launch (the non-deferred form)
        dg = kr[f"site_dgrad_grouped_M{2 * Mh}"]
        wg = kr["wgrad_deferred_12sites"]
        fwd = kr[f"site_fwd_grouped_M{2 * Mh}"]
        big = kr[f"bwd_gating_M{12 * Mh}"]
        alg_bytes, alg_flops = 6 * D * 2 * Mh, 20 * D * RANK * Mh
        bwd_us = round(dg["us"] + wg["us_per_site"], 2)
        bwd_gbs = round(alg_bytes / (bwd_us * 1e-6) / 1e9, 1)
        bwd_tf = round(alg_flops / (bwd_us * 1e-6) / 1e12, 1)
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "r2_dat_traffic.json"
        if tp.exists():
            tj = json.loads(tp.read_text())
            k = tj["kernels"]
            if "site_dgrad_grouped" in k and "wgrad_deferred_12sites" in k:
                traffic = (k["site_dgrad_grouped"]["dram_read_bytes"] + k["site_dgrad_grouped"]["dram_write_bytes"]
                           + (k["wgrad_deferred_12sites"]["dram_read_bytes"] + k["wgrad_deferred_12sites"]["dram_write_bytes"]) / 12)
                traffic_src = "profiles/r2_dat_traffic.json (" + tj["source"] + ")"
        # share of the DAT kernels in the step, from the launches the step makes (12 x fwd, 12 x dgrad, 1 x wgrad)
        dat_us = 12 * (fwd["us"] + dg["us"]) + wg["us"]
        result = {
            "metric": METRIC, "value": round(value, 2), "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "clients": world,
                       "step_launch": "eager" if args.eager else "cuda-graph replay of train_step",
                       "round_boundary": f"one FedAvg allreduce of the flat adapter_1 buffer after the {K} timed steps "
                                         "(inside the timed region, communicator warmed; allreduce_us = its own time)",
                       "l2": "per-step working set (222 MB bf16 backbone weights + >1 GB activations) exceeds the 126 MB L2; kernel micro-timings rotate through input sets totalling > 2x L2 (no flush writes)",
                       "init": "seeded random ViLT-B/32 (no pretrained weights on the box)"},
            "clocks": clocks,
            "e2e": {"value": round(e2e, 2), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": round(ms_e2e / K, 3)},
            "gpu_launches": launches,
            "allreduce_us": allreduce_us,
            "roofline": {"kernel": "DAT backward of one adapter site as the batched MKD schedule launches it: ONE grouped dgrad "
                                   f"launch over {Mh} gating rows (R = {2 * RANK}) + {Mh} adapter_1 rows (R = {RANK}), plus the "
                                   "site's 1/12 share of the ONE deferred wgrad launch over all 12 sites",
                         "bound": "hbm", "achieved": bwd_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(bwd_gbs / peaks["hbm_gbs"], 4),
                         "traffic": traffic,
                         "traffic_unit": "dram read + write bytes: the site's dgrad launch + 1/12 of the 12-site wgrad launch (ncu --set full)",
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_site": alg_bytes, "algorithmic_flops_per_site": alg_flops,
                         "tensor_view": {"achieved_tflops": bwd_tf, "peak": peaks["tf_burst"],
                                         "frac": round(bwd_tf / peaks["tf_burst"], 4)},
                         "peak_source": f"{peaks['source']} hbm_gbs / bf16_tflops burst (kernels timed alone, CUDA events, cold inputs)",
                         "us": bwd_us,
                         "parts": {"dgrad_launch_us": dg["us"], "wgrad_12site_launch_us": wg["us"],
                                   "wgrad_launch_gbs": wg["gbs"], "wgrad_launch_frac": wg["frac_of_roofline"]},
                         "per_site_launch_pair_us": pair["us"],
                         "forward_same_site": {"us": fwd["us"], "achieved": fwd["gbs"], "frac": fwd["frac_of_roofline"]},
                         "fwd_plus_bwd_site": {"us": round(fwd["us"] + bwd_us, 2),
                                               "frac": round((10 * D * 2 * Mh) / ((fwd["us"] + bwd_us) * 1e-6) / 1e9 / peaks["hbm_gbs"], 4)},
                         "dat_us_per_step_from_these": round(dat_us, 1),
                         "steady_state_M71040": {"kernel": "bwd gating, single group (dat_pipe_kernel + wgrad)",
                                                 "achieved_tflops": big["tflops"], "frac": round(big["tflops"] / peaks["tf_burst"], 4)}},
            "kernels": kr,
        }
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_reference(steps=2, warmup=1, sample_batch=args.cpu_batch)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result


# ------------------------------------------------------------------------------------------------
ALBEF_WORKLOAD = ("ALBEF (ViT-B/16 @384 + 12-layer BERT question encoder + 6-layer LM-head answer decoder) + DAT rank-256 "
                  "bf16, 30 adapter sites, synthetic VQA batch=16 (32 answers x 6 tokens, vocabulary 30522), MKD tau=2.0 "
                  "(BASELINE configs[2])")


def build_albef_client(rank: int, device, batch: int = 16, rank_r: int = 256):
    """Model + trainer of one ALBEF client (BASELINE configs[2]: full depth, rank 256, random init, seeded)."""
    import torch
    from feddat_b200.modeling.albef import convert_batch_to_albef_input_dict
    from feddat_b200.train.accelerator import Accelerator
    from feddat_b200.train.prepare import default_args, place_on_gpu, prepare_model
    from feddat_b200.train.task_trainer import TaskTrainer, get_polynomial_decay_schedule_with_warmup
    torch.manual_seed(2000)
    a = default_args(encoder_name="albef_no_distill", ordered_cl_tasks=[f"synth{rank % 8}"], adapter_rank=rank_r, image_size=384)
    model = prepare_model(a, place=False)
    sd = model.state_dict()
    for name in sd:
        if "adapter_1" in name:
            sd[name.replace("adapter_1", "adapter_2")].data.copy_(sd[name].data)
    for n, p in model.named_parameters():
        if "adapter_2" in n:
            p.requires_grad = False
    place_on_gpu(model, device)
    tr = TaskTrainer()
    tr.args = SimpleNamespace(optimizer_mode="dat", encoder_name="albef_no_distill", debug=0)
    tr.accelerator = Accelerator(device=device)
    tr.device, tr.task_key = torch.device(device), a.ordered_cl_tasks[0]
    tr.batch2inputs_converter = convert_batch_to_albef_input_dict
    tr.weight_decay, tr.lr, tr.adam_epsilon, tr.kl_temp = 1e-2, 1e-4, 1e-8, TEMP
    wrapped = tr.accelerator.prepare(model)
    opt = tr.create_optimizer(wrapped)
    sched = get_polynomial_decay_schedule_with_warmup(opt, 100, 100000, lr_end=0, power=1)
    wrapped.train()
    return tr, wrapped, opt, sched


def run_albef(args):
    """``--workload albef``: one TaskTrainer.train_step of the ALBEF path (three forwards, two backwards, two AdamW
    steps in the reference order -- BERT has dropout, so no pass is shared) per step, replayed from a CUDA graph.
    An extra line beside the headline ViLT workload; same timing rules."""
    import torch
    import torch.distributed as dist
    from feddat_b200 import ops
    from feddat_b200.synthetic import albef_to_device, make_albef_batch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    BA, RA = 16, 256
    tr, wrapped, opt, sched = build_albef_client(rank, device, BA, RA)
    K, W = args.steps, args.warmup
    host = [make_albef_batch(BA, 384, seed=(2000 + rank) * 1000 + i, client=rank % 8, pin=True) for i in range(min(K + W, 6))]
    devb = [albef_to_device(b, device) for b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    if args.eager:
        step = lambda b_: tr.train_step(wrapped, 0, b_, opt, sched)                      # noqa: E731
        step_e2e = lambda hb: tr.train_step(wrapped, 0, albef_to_device(hb, device), opt, sched)   # noqa: E731
    else:
        from feddat_b200.train.graphed import GraphedDictStep
        graphed = GraphedDictStep(tr, wrapped, opt, sched, devb[0], warmup=2)
        step = graphed                    # device batch -> static tensors -> graph replay
        # pinned host batch -> device (H2D inside the timed region; the fp32 -> bf16 image cast runs on the GPU) ->
        # static tensors -> replay
        step_e2e = lambda hb: graphed(albef_to_device(hb, device))                        # noqa: E731
    for i in range(max(W, 3)):
        step(devb[i % len(devb)])
    l0 = ops.launch_count
    ms = timed(lambda i: step(devb[i % len(devb)]), K)
    launches = ops.launch_count - l0
    ms_e2e = timed(lambda i: step_e2e(host[i % len(host)]).item(), K)
    res = None
    if rank == 0:
        h2d = sum(v.numel() * v.element_size() for v in host[0].values() if hasattr(v, "numel"))
        # the fused MKD head alone, at this workload's logits: 32 answers x 6 tokens x 30522 bf16
        g = torch.Generator(device=device).manual_seed(0)
        n_seq, La, Cv = sum(host[0]["n"]), host[0]["answer_ids"].shape[1], 30522
        sets = [(torch.randn(n_seq, La, Cv, device=device, generator=g).to(torch.bfloat16),
                 torch.randn(n_seq, La, Cv, device=device, generator=g).to(torch.bfloat16)) for _ in range(24)]
        lab = devb[0]["answer_ids"].masked_fill(devb[0]["answer_ids"] == 0, -100)
        sw = devb[0]["weights"] / BA
        ts = []
        for i in range(3 + 12):
            sc, te = sets[i % len(sets)]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(500_000)
            e0.record()
            ops.mkd_ce_loss(sc, te[:, :-1], lab, sw, TEMP)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e-3)
        t_head = statistics.mean(ts)
        head_bytes = 3 * n_seq * La * Cv * 2
        peaks = read_peaks()
        res = {"metric": "VQA samples/sec/box (ALBEF+DAT bf16)", "value": round(world * BA * K / (ms * 1e-3), 2),
               "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms / K, 3),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": ALBEF_WORKLOAD, "global_batch": BA * world, "clients": world,
                          "step_launch": "eager" if args.eager else "cuda-graph replay of train_step", "init": "seeded random ALBEF (no checkpoint on the box)"},
               "e2e": {"value": round(world * BA * K / (ms_e2e * 1e-3), 2), "unit": "samples/s",
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / K, 3)},
               "gpu_launches": launches,
               "kernels": {"mkd_ce_loss": {"us": round(t_head * 1e6, 2), "rows": n_seq * La, "C": Cv, "dtype": "bf16",
                                           "algorithmic_bytes": head_bytes,
                                           "gbs": round(head_bytes / t_head / 1e9, 1),
                                           "frac_of_hbm_peak": round(head_bytes / t_head / 1e9 / peaks["hbm_gbs"], 4)}}}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return res


def cpu_reference(steps: int, warmup: int, sample_batch: int):
    """The reference algorithm (oracle/step_oracle.py, a PyTorch-CPU port pinned to the reference's
    own trainer by tests/test_step_oracle.py) on all host cores: fp32, same model / image / text
    sizes, a bounded batch.  /root/reference itself cannot travel to the GPU box."""
    import torch
    from feddat_b200.synthetic import make_vilt_batch
    from oracle import step_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1000)
    model = step_oracle.OracleLearner(RANK, tasks=("synth0",), num_labels=C).prepare_dat()
    opt = step_oracle.create_optimizer(model, 1e-4)
    sched = step_oracle.create_scheduler(opt, 1000)
    model.train()
    times = []
    for i in range(warmup + steps):
        batch = make_vilt_batch(sample_batch, T, H, C, seed=1000000 + i)
        t0 = time.time()
        step_oracle.train_step(model, "synth0", batch, opt, sched, temp=TEMP)
        if i >= warmup:
            times.append(time.time() - t0)
    sec = statistics.median(times)
    return {"value": round(sample_batch / sec, 3), "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{steps} train_step(s) (+{warmup} warm-up) of batch {sample_batch} (the workload's is {B}) at 384x384 / 40 tok, "
                      f"rank {RANK}, fp32, torch CPU {cores} threads; median {sec:.2f} s/step",
            "ms_per_step": round(sec * 1e3, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    K, W = args.steps, args.warmup
    t0 = time.time()
    base = cpu_reference(steps=K, warmup=W, sample_batch=args.cpu_batch)
    return {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": args.cpu_batch, "clients": 1, "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.time() - t0, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=None,
                    help="batch of the CPU arm (default: the workload's own 32, so both arms run the same config; "
                         "4 when more than 30 CPU steps are requested, to stay within minutes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch train_step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--workload", default="vilt", choices=["vilt", "albef"],
                    help="vilt = BASELINE configs[1] (the headline metric); albef = configs[2], an extra line")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.cpu_batch is None:
        args.cpu_batch = B if (args.impl != "reference" or args.steps + args.warmup <= 30) else 4
    if args.workload == "albef" and args.impl == "ours":
        res = run_albef(args)
        if res is not None:
            print(json.dumps(res), flush=True)
        return
    res = run_reference(args) if args.impl == "reference" else run_ours(args)
    if res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()

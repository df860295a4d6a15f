"""ncu raw page (``ncu -i X.ncu-rep --page raw --csv``) of scripts/profile_grouped.py -> profiles/r2_dat_traffic.json:
per-launch dram__bytes_read.sum + dram__bytes_write.sum and duration of the grouped DAT launches of one site and of the deferred
weight-gradient launch over all 12 sites (the LAST launch of each kernel in the capture).  bench.py reads the JSON for ``roofline.traffic``.
    python scripts/ncu_traffic.py gpurun_out/r2_grouped_raw.csv"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src = Path(sys.argv[1])
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3}
names = {"dat_fused_kernel<0, 0, 0>": "site_fwd_grouped", "dat_fused_kernel<1, 0, 1>": "site_dgrad_grouped",
         "dat_wgrad_kernel": "wgrad_deferred_12sites", "pack_kernel": "pack_batched"}


def val(r, k):
    return float(r[idx[k]].replace(",", "")) * mul[units[idx[k]]]


out = {"source": f"{src.name}: ncu --set full --clock-control none, scripts/profile_grouped.py (one site's inputs evicted "
                 "from L2 by a 256 MB memset before its forward / data-gradient launch; the weight-gradient launch reads the "
                 "0.5 GB of all 12 sites)", "kernels": {}}
for r in rows[2:]:
    for pat, key in names.items():
        if pat in r[idx["Kernel Name"]]:
            out["kernels"][key] = {
                "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                "ncu_duration_us": val(r, "gpu__time_duration.sum"),
                "tensor_pipe_active_pct": float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
                "grid": int(r[idx["launch__grid_size"]])}
(ROOT / "profiles" / "r2_dat_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))

"""Opcode histogram per kernel of the shipped library -> profiles/r2_sass_opcodes.txt
    python scripts/sass_histogram.py
(cuobjdump -sass; tensor-core / TMA / TMEM / mbarrier / packed-fp32 opcodes are listed with their modifiers, the rest as
the twelve most frequent base opcodes)."""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "feddat_b200" / "lib" / "libfeddat_sm100.so"
KEEP = ("UTC", "UTMA", "LDTM", "STTM", "SYNCS", "UBLKCP", "ACQBULK", "MUFU", "FFMA2", "FMUL2", "FADD2", "UTCBAR", "REDG", "RED.")
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
kern, full, base = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        full[kern], base[kern] = collections.Counter(), collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        base[kern][op.split(".")[0]] += 1
        if any(op.startswith(k) for k in KEEP):
            op = re.sub(r"\.(64|128|32|16)$", "", op)
            full[kern][".".join(op.split(".")[:4])] += 1
lines = [f"opcode histogram per kernel of feddat_b200/lib/{LIB.name} (cuobjdump -sass; tensor / TMA / TMEM / packed-fp32 "
         "opcodes kept with their modifiers); scripts/sass_histogram.py", ""]
for k in full:
    lines.append(k)
    lines.append(f"  total {sum(base[k].values())}  |  " + ", ".join(f"{o} {n}" for o, n in sorted(full[k].items())))
    lines.append("  top: " + ", ".join(f"{o} {n}" for o, n in base[k].most_common(12)))
    lines.append("")
(ROOT / "profiles" / "r2_sass_opcodes.txt").write_text("\n".join(lines))
print(f"{len(full)} kernels")

"""profiles/resource_usage.md: registers / stack / static smem per kernel of the built extension
(`cuobjdump -res-usage`, CPU box).  LOCAL:0 everywhere = no register spills."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "zero_transformer_b200", "_ext", "_ztb_cuda.so")
txt = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
rows, cur = [], None
for line in txt.split("\n"):
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        rows.append((cur, *map(int, m.groups())))
        cur = None
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
out = ["# Kernel resource usage", "",
       "`python bench/res_usage.py` = `cuobjdump -res-usage` on the built extension (sm_100a).  Dynamic shared memory (the TMA rings:",
       "~225 KB for the GEMM and attention kernels) is not listed by this tool; `LOCAL` = spill bytes, `STACK` = the by-value peer-pointer tables.", "",
       "| kernel | registers | stack B | static smem B | local (spill) B |", "|---|---|---|---|---|"]
for (mangled, reg, stack, sh, loc), n in zip(rows, names):
    out.append(f"| `{n.strip()[:100]}` | {reg} | {stack} | {sh} | {loc} |")
open(os.path.join(ROOT, "profiles", "resource_usage.md"), "w").write("\n".join(out) + "\n")
print("kernels:", len(rows), "max spill:", max(r[4] for r in rows))

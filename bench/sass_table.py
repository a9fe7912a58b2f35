"""profiles/sass_evidence.md: per-kernel counts of the Blackwell-specific SASS mnemonics in the built extension
(`cuobjdump -sass`, runs on the CPU box)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "zero_transformer_b200", "_ext", "_ztb_cuda.so")
COLS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
        ("UTMAREDG", r"\bUTMAREDG"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("LDGMC", r"\bLDGMC"), ("STS.128", r"\bSTS\.128"),
        ("HMMA", r"\bHMMA")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, order, cur, i = collections.defaultdict(collections.Counter), [], None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[i].strip()
            i += 1
            order.append(cur)
            continue
        if cur is None:
            continue
        for label, pat in COLS:
            if re.search(pat, line):
                counts[cur][label] += 1
    out = ["# SASS evidence", "",
           "`python bench/sass_table.py` = `cuobjdump -sass zero_transformer_b200/_ext/_ztb_cuda.so`, counts of Blackwell-specific mnemonics per kernel.",
           "`UTC*MMA` = `tcgen05.mma`, `LDTM`/`STTM` = `tcgen05.ld`/`st`, `UTMALDG`/`UTMASTG`/`UTMAREDG` = TMA tensor load / store / reduce-add,",
           "`UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier ops, `LDGMC` = `multimem.ld_reduce` (NVLS switch-side reduction; the matching",
           "`multimem.st` is an `STG...STRONG.SYS` to the multicast address).  No `HMMA` (legacy mma.sync) appears in any kernel.", "",
           "| kernel | " + " | ".join(c for c, _ in COLS) + " |", "|---|" + "---|" * len(COLS)]
    for k in order:
        c = counts[k]
        if not any(c[x] for x, _ in COLS):
            continue
        out.append(f"| `{k[:96]}` | " + " | ".join(str(c[x]) for x, _ in COLS) + " |")
    path = os.path.join(ROOT, "profiles", "sass_evidence.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path, len(order), "kernels")


if __name__ == "__main__":
    sys.exit(main())

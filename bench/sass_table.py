"""profiles/sass_evidence.md: per-kernel counts of the Blackwell-specific SASS mnemonics in the built extension
(`cuobjdump -sass`, runs on the CPU box)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "zero_transformer_b200", "_ext", "_ztb_cuda.so")
HOT = ["gemm2_tcgen05_kernel<false, false, 0, false>", "gemm2_tcgen05_kernel<false, false, 1, false>", "gemm2_tcgen05_kernel<false, false, 0, true>",
       "gemm2_tcgen05_kernel<true, true, 0, false>", "gemm_tcgen05_kernel<false, false, false>", "attn_fwd3_kernel", "attn_bwd_kernel<128>",
       "adamw_shard_kernel<false, false>", "layernorm_bwd_kernel<8>"]
KEY = re.compile(r"UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTCBAR|UBLKCP|UBLKRED|SYNCS|LDGMC|MULTIMEM|\.STRONG\.SYS|UTCATOMSWS|UCGABAR|ACQBULK|RED\.")
COLS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
        ("UTMAREDG", r"\bUTMAREDG"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("LDGMC", r"\bLDGMC"), ("UBLKCP", r"\bUBLKCP"), ("STS.128", r"\bSTS\.128"),
        ("HMMA", r"\bHMMA")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, order, cur, i = collections.defaultdict(collections.Counter), [], None, 0
    body = collections.defaultdict(list)
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[i].strip()
            i += 1
            order.append(cur)
            continue
        if cur is None:
            continue
        if re.search(r"/\*[0-9a-f]{4}\*/", line):
            body[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
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
    # raw excerpts: for every hot kernel, each instruction that is a tensor-core / TMA / bulk-copy / mbarrier / multimem op,
    # with one line of context before it, straight from cuobjdump (addresses kept)
    sdir = os.path.join(ROOT, "profiles", "sass")
    os.makedirs(sdir, exist_ok=True)
    for want in HOT:
        ks = [k for k in order if want in k]
        if not ks:
            continue
        k = ks[0]
        lines, keep = body[k], set()
        for j, ln in enumerate(lines):
            if KEY.search(ln):
                keep.update((j - 1, j))
        fn = re.sub(r"[^A-Za-z0-9]+", "_", want).strip("_") + ".sass.txt"
        with open(os.path.join(sdir, fn), "w") as f:
            f.write(f"// cuobjdump -sass _ztb_cuda.so :: {k}\n// {len(lines)} instructions; excerpt = tensor-core / TMA / bulk / mbarrier / multimem instructions (+1 line of context)\n")
            last = -2
            for j in sorted(x for x in keep if x >= 0):
                if j != last + 1:
                    f.write("        ...\n")
                f.write(lines[j] + "\n")
                last = j
    path = os.path.join(ROOT, "profiles", "sass_evidence.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path, len(order), "kernels")


if __name__ == "__main__":
    sys.exit(main())

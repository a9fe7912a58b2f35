"""Turn gpurun_out/*.ncu-rep into committed markdown summaries under profiles/ (run on the CPU box)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__cycles_active.avg", "SM active cycles"),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    reps = sys.argv[1:]
    for rep in reps:
        hdr, units, rows = load(rep)
        name = os.path.splitext(os.path.basename(rep))[0]
        lines = [f"# ncu summary: {name}", "",
                 f"Source: `{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`, 1x B200). "
                 f"Roofline denominators (measured): HBM {PEAKS.get('hbm_gbs')} GB/s, bf16 {PEAKS.get('bf16_tflops')} TFLOP/s burst / "
                 f"{PEAKS.get('bf16_tflops_sustained', 'n/a')} sustained. Durations under ncu are for shares only; timings quoted in DESIGN.md are CUDA-event numbers.", ""]
        ik = hdr.index("Kernel Name")
        for r in rows:
            lines.append(f"## `{r[ik][:110]}`")
            lines.append("")
            lines.append("| metric | value |")
            lines.append("|---|---|")
            vals = {}
            for key, label in METRICS:
                if key in hdr:
                    i = hdr.index(key)
                    vals[key] = (r[i], units[i])
                    lines.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
            try:
                dur_us = float(vals["gpu__time_duration.sum"][0].replace(",", ""))
                u = vals["gpu__time_duration.sum"][1]
                dur_s = dur_us * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1e-6)
                tot = 0.0
                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, uu = vals[k]
                    tot += float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(uu, 1.0)
                lines.append(f"| DRAM traffic / duration | {tot / dur_s / 1e9:.0f} GB/s = {100 * tot / dur_s / 1e9 / PEAKS.get('hbm_gbs', 6650):.1f} % of measured copy bandwidth |")
            except Exception:
                pass
            lines.append("")
        os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
        with open(os.path.join(ROOT, "profiles", name + ".md"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("wrote profiles/" + name + ".md")


if __name__ == "__main__":
    main()

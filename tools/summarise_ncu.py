"""One markdown table per `ncu --set full` report: every captured launch with its duration, DRAM traffic, achieved DRAM
bandwidth against the measured HBM peak, pipe utilisation and launch geometry (development tool).

    python tools/summarise_ncu.py gpurun_out/r02_prof_hbm.ncu-rep profiles/r02_hbm_kernels_ncu.md "title" "command"
Optionally a JSON file (5th argument) receives {"dram_bytes_per_launch_mean": ...} over the launches whose kernel name
matches the 6th argument (used for bench.py's roofline.traffic)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0,
        "hz": 1.0, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9}


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def short(name):
    name = name.replace("pdm::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_]+)(<[^>(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    hdr, units, rows = load(rep)

    def val(r, k, scale=True):
        if k not in hdr:
            return float("nan")
        i = hdr.index(k)
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            return float("nan")
        return v * UNIT.get(units[i], 1.0) if scale else v

    lines = [f"# {title}\n", f"Command (after the same command exited 0 without ncu): `{cmd}`\n" if cmd else "",
             f"DRAM bandwidth = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum, against the measured copy peak "
             f"{hbm} GB/s (`MEASURED_PEAKS.json`).  ncu runs every kernel alone, cold-cache, at its own clocks.\n",
             "| # | kernel | grid x block | regs | duration us | dram read MB | dram write MB | DRAM GB/s | % of HBM peak | tensor pipe % | XU pipe % | issue active % | L2 hit % | SM GHz |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic = []
    for i, r in enumerate(rows):
        name = short(r[hdr.index("Kernel Name")])
        dur = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        bw = (rd + wr) / dur / 1e9 if dur > 0 else float("nan")
        lines.append("| {} | `{}` | {} x {} | {} | {:.1f} | {:.2f} | {:.2f} | {:.0f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.2f} |".format(
            i, name, r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "?",
            r[hdr.index("launch__block_size")] if "launch__block_size" in hdr else "?",
            r[hdr.index("launch__registers_per_thread")] if "launch__registers_per_thread" in hdr else "?",
            dur * 1e6, rd / 1e6, wr / 1e6, bw, 100 * bw / hbm,
            val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
            val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", False),
            val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
            val(r, "lts__t_sector_hit_rate.pct", False),
            val(r, "sm__cycles_elapsed.max.per_second") / 1e9))
        traffic.append((name, rd + wr))
    open(out_md, "w").write("\n".join(l for l in lines if l is not None) + "\n")
    if len(sys.argv) > 6:
        sel = [t for n, t in traffic if re.search(sys.argv[6], n)]
        if sel:
            json.dump({"dram_bytes_per_launch_mean": sum(sel) / len(sel), "launches": len(sel), "report": os.path.basename(rep)},
                      open(sys.argv[5], "w"))
    print(f"{out_md}: {len(rows)} launches")


if __name__ == "__main__":
    main()

"""Turn the ncu captures under gpurun_out/ into the tracked summaries under profiles/ (development tool).
   python tools/summarise_profiles.py r01b
Inputs:  gpurun_out/launches_<tag>.csv            (ncu --metrics gpu__time_duration.sum launch list of bench.py)
         gpurun_out/prof_bench_gemm_<tag>.ncu-rep  (ncu --set full, -k regex:gemm_tc, inside bench.py)
         gpurun_out/prof_bench_attn_<tag>.ncu-rep  (ncu --set full, -k regex:attention_tc3, inside bench.py)"""
import csv, io, json, os, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01b"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launches():
    src = os.path.join(G, f"launches_{tag}.csv")
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("pdm::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e3  # ns -> us
    tot = sum(v[1] for v in agg.values())
    out = [f"# r01 -- ncu launch list of `python bench.py --steps 1 --warmup 3` ({sum(v[0] for v in agg.values())} launches after skipping 30 000)\n",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 600 --csv ... python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile`",
           "(per-launch times are cold-cache and serialised: compare SHARES with `bench.py`'s `kernels` object, not absolutes)\n",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |")
    open(os.path.join(P, "r01_bench_launches_summary.md"), "w").write("\n".join(out) + "\n")
    import shutil
    shutil.copy(src, os.path.join(P, "r01_bench_launches.csv"))


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[2:]


def table(rep, title, cmd, fname, note):
    hdr, rows = raw(rep)
    g = lambda r, k: r[hdr.index(k)] if k in hdr else "n/a"
    out = [f"# r01 -- `ncu --set full` capture of `{title}` inside `bench.py`\n", f"Command (after the same command exited 0 without ncu): `{cmd}`\n",
           "| launch | duration us | dram read MB | dram write MB | tensor pipe active % | XU pipe % | issue active % | dram % of peak | regs | grid | SM GHz |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic = []
    for i, r in enumerate(rows):
        rd = float(g(r, "dram__bytes_read.sum")); wr = float(g(r, "dram__bytes_write.sum"))
        ur, uw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        out.append("| {} | {:.1f} | {} | {} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {} | {} | {:.2f} |".format(
            i, float(g(r, "gpu__time_duration.sum")) , g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum"),
            float(g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")),
            float(g(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")),
            float(g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")),
            float(g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
            g(r, "launch__registers_per_thread"), g(r, "launch__grid_size"), float(g(r, "sm__cycles_elapsed.max.per_second"))))
        traffic.append((rd, wr))
    out.append("\n(dram columns are in the units ncu printed for that launch: see the raw page of the report; durations in the unit of `gpu__time_duration.sum` of that row)\n")
    out.append(note)
    open(os.path.join(P, fname), "w").write("\n".join(out) + "\n")
    return hdr, rows


if __name__ == "__main__":
    launches()
    print("launch summary written")

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
    out = [f"# {tag} -- ncu launch list of `python bench.py --steps 1 --warmup 3` ({sum(v[0] for v in agg.values())} launches after skipping 30 000)\n",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 700 --csv ... python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile`",
           "(per-launch times are cold-cache and serialised: compare SHARES with `bench.py`'s `kernels` object, not absolutes)\n",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |")
    open(os.path.join(P, f"{tag}_bench_launches_summary.md"), "w").write("\n".join(out) + "\n")
    import shutil
    shutil.copy(src, os.path.join(P, f"{tag}_bench_launches.csv"))


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


GEMM_LABELS = {  # launch order of `-k regex:gemm_tc -s 3000 -c 8` in the two-stream small config (deferred-LayerNorm flow)
    "r01d": ["gemm_proj  mask stream  M=302080 N=K=512, fp32 RMW + bf16 copy + LN row sums, TMA-in / TMA-out epilogue   <2,F32_TMA>",
             "gemm_fc1   mask stream  M=302080 N=2048 K=512, LN folded + GELU, bf16 out, 16 epilogue warps   <2,LN_GELU_W16>",
             "gemm_fc2   mask stream  M=302080 N=512 K=2048, fp32 RMW + 2 bf16 copies + LN row sums   <2,EMIT>",
             "gemm_zeroconv  M=512x334 rows, N=K=512, fp32 RMW + fp32/bf16 concat copies + LN row sums (x2), TMA epilogue   <2,F32_TMA>",
             "gemm_qkv   image stream M=171008 N=1536 K=512, LN folded, bf16 out   <2,LN>",
             "gemm_proj  image stream M=171008 N=K=512   <2,F32_TMA>",
             "gemm_fc1   image stream M=171008 N=2048 K=512   <2,LN_GELU_W16>",
             "gemm_fc2   image stream M=171008 N=512 K=2048, fp32 RMW only (zero-conv follows)   <2,F32>"],
    "r01c": ["gemm_proj  mask stream  M=302080 N=K=512, fp32 RMW + bf16 copy + LN row sums   <2,EMIT>",
             "gemm_fc1   mask stream  M=302080 N=2048 K=512, LN folded + GELU, bf16 out, 16 epilogue warps   <2,LN_GELU_W16>",
             "gemm_fc2   mask stream  M=302080 N=512 K=2048, fp32 RMW + 2 bf16 copies + LN row sums   <2,EMIT>",
             "gemm_zeroconv  M=512x334 rows, N=K=512, fp32 RMW + fp32/bf16 concat copies + LN row sums (x2)   <2,EMIT>",
             "gemm_qkv   image stream M=171008 N=1536 K=512, LN folded, bf16 out   <2,LN>",
             "gemm_proj  image stream M=171008 N=K=512   <2,EMIT>",
             "gemm_fc1   image stream M=171008 N=2048 K=512   <2,LN_GELU_W16>",
             "gemm_fc2   image stream M=171008 N=512 K=2048, fp32 RMW only (zero-conv follows)   <2,F32>"],
}


def full(kind, rep, title, cmd, fname, note, labels=None):
    hdr, rows = raw(rep)
    g = lambda r, k: r[hdr.index(k)] if k in hdr else "nan"
    out = [f"# {tag} -- `ncu --set full` capture of `{title}` inside `bench.py`\n",
           f"Command (after the same command exited 0 without ncu): `{cmd}`\n",
           "| launch | duration us | dram read GB | dram write GB | tensor pipe active % | XU (MUFU) pipe % | issue active % | dram % of peak | L2 % of peak | regs | grid | SM GHz |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    tot = []
    for i, r in enumerate(rows):
        unit = lambda k: {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[rows_units[hdr.index(k)]]
        rd, wr = float(g(r, "dram__bytes_read.sum")) * unit("dram__bytes_read.sum"), float(g(r, "dram__bytes_write.sum")) * unit("dram__bytes_write.sum")
        tot.append((rd + wr) * 1e9)
        out.append("| {} | {:.1f} | {:.3f} | {:.3f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {} | {} | {:.2f} |".format(
            labels[i] if labels and i < len(labels) else i, float(g(r, "gpu__time_duration.sum")), rd, wr,
            float(g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")),
            float(g(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")),
            float(g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")),
            float(g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
            float(g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")),
            g(r, "launch__registers_per_thread"), g(r, "launch__grid_size"), float(g(r, "sm__cycles_elapsed.max.per_second"))))
    out.append("\n" + note)
    open(os.path.join(P, fname), "w").write("\n".join(out) + "\n")
    return tot


if __name__ == "__main__":
    launches()
    print("launch summary written")
    rows_units = None
    for kind in ("gemm", "attn"):
        rep = os.path.join(G, f"prof_bench_{kind}_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows_units = list(csv.reader(io.StringIO(txt)))[1]
        if kind == "gemm":
            note = open(os.path.join(P, f"{tag}_gemm_note.txt")).read() if os.path.exists(os.path.join(P, f"{tag}_gemm_note.txt")) else ""
            tot = full(kind, rep, "gemm_tc_kernel<2, EPI>",
                       "ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3000 -c 8 -o prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-profile",
                       f"{tag}_gemm_ncu_summary.md", note, GEMM_LABELS.get(tag))
            json.dump({"kernel": "gemm_tc_kernel<2, EPI>", "source": f"profiles/{tag}_gemm_ncu_summary.md",
                       "dram_bytes_per_launch_mean": sum(tot) / len(tot), "launches_captured": len(tot)},
                      open(os.path.join(P, f"{tag}_gemm_traffic.json"), "w"), indent=1)
        else:
            full(kind, rep, "attention_tc3_kernel",
                 "ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 600 -c 2 -o prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-profile",
                 f"{tag}_attention_ncu_summary.md", "Launch 0: image stream (L = 334), launch 1: mask stream (L = 590); nb = 512, H = 8.")
    print("ncu summaries written")

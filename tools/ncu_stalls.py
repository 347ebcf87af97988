"""Summarise an .ncu-rep: key metrics + top stall instructions (development tool).
   python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:80])
    for k in keys:
        if k in hdr:
            print(f"  {k:75s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# may contain several kernels; take the first table
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
data = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    data.append(r)
isamp, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
agg = {s: sum(int(r[hdr.index(s)] or 0) for r in data) for s in stalls}
print("total samples", tot, "instructions", len(data))
print({k[6:]: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.01 * tot})
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:ntop]
for i in sorted(top):
    r = data[i]
    st = {s[6:]: int(r[hdr.index(s)] or 0) for s in stalls}
    st = {k: v for k, v in st.items() if v > 0.15 * int(r[isamp])}
    print(f"{i:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>9s} {r[isrc][:80]:80s} {st}")

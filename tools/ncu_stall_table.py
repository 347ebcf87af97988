"""Warp-stall breakdown of the launches in an `ncu --set full --import-source on` report, per kernel launch and per opcode
(development tool; the source page attributes every sample to a SASS instruction).

    python tools/ncu_stall_table.py prof.ncu-rep out.md "title" [max_launches]"""
import collections
import csv
import io
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
maxl = int(sys.argv[4]) if len(sys.argv) > 4 else 99
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
names = [rows[i - 1][1] if i > 0 and len(rows[i - 1]) > 1 else "" for i in starts]
out = [f"# {title}\n", "Per launch: share of warp-stall samples by reason, and the opcodes that hold most samples "
       "(reasons holding >= 15 % of that opcode's samples).\n"]
for li, st in enumerate(starts[:maxl]):
    hdr = rows[st]
    data = []
    for r in rows[st + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        data.append(r)
    isamp, isrc = hdr.index("# Samples"), hdr.index("Source")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp] or 0) for r in data) or 1
    agg = collections.Counter()
    byop = collections.defaultdict(collections.Counter)
    for r in data:
        op = r[isrc].split()[1] if r[isrc].startswith("@") else (r[isrc].split()[0] if r[isrc].split() else "?")
        op = op.split(".")[0]
        byop[op]["samples"] += int(r[isamp] or 0)
        for s in stalls:
            v = int(r[hdr.index(s)] or 0)
            agg[s[6:]] += v
            byop[op][s[6:]] += v
    out.append(f"## launch {li}: {names[li][:120]}\n")
    out.append("stall reasons: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in agg.most_common(9)) + "\n")
    out.append("| opcode | samples | share | dominant reasons |")
    out.append("|---|---|---|---|")
    for op, c in sorted(byop.items(), key=lambda kv: -kv[1]["samples"])[:10]:
        top = ", ".join(f"{k} {100 * v / max(c['samples'], 1):.0f} %" for k, v in c.most_common(6) if k != "samples" and v >= 0.15 * c["samples"])
        out.append(f"| {op} | {c['samples']} | {100 * c['samples'] / tot:.1f} % | {top} |")
    out.append("")
open(out_md, "w").write("\n".join(out) + "\n")
print(out_md, len(starts), "launches")

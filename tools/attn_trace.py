"""Decode the per-warp event trace written by a -DPDM_ATTN_TRACE build (development tool)."""
import sys, struct
import numpy as np
d = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(12, 4096)
names = {0: "wait_S", 1: "got_S", 2: "exp_done", 3: "arrived", 6: "S_in_regs", 7: "max_done", 8: "got_O", 4: "epi_start", 5: "epi_done", 10: "QK_issued", 11: "wait_P", 12: "got_P", 13: "PV_issued", 14: "wait_free", 15: "got_free"}
t0 = min(int(x & 0xffffffffff) for w in range(12) for x in d[w] if x)
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 60
for w in (0, 4, 9, 10):
    print("warp", w)
    prev = None
    for x in d[w][lo:hi]:
        x = int(x)
        if not x: break
        ev, n, c = x >> 56, (x >> 40) & 0xffff, (x & 0xffffffffff) - t0
        print(f"   {names.get(ev, ev):10s} n={n:4d} t={c:8d}" + (f"  (+{c - prev})" if prev is not None else ""))
        prev = c

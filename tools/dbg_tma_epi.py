"""Development aid: layout check of the TMA-store GEMM epilogue (resid pattern in -> same pattern out)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panopticdiffusionmodels_b200 import _lib
dev = torch.device("cuda:0")
M, N, K = 256, 256, 64
A = torch.zeros(M, K, device=dev)
W = torch.zeros(N, K, device=dev)
bias = torch.zeros(N, device=dev)
r = torch.arange(M, device=dev).float().view(M, 1) * 1000 + torch.arange(N, device=dev).float().view(1, N)
out = torch.empty(M, N, device=dev)
_lib.check(_lib.lib().pdm_debug_linear(_lib.ptr(A), None, _lib.ptr(W), _lib.ptr(bias), _lib.ptr(r), _lib.ptr(out), M, N, K, 0, 0, 0, 0, None, _lib.current_stream()))
torch.cuda.synchronize()
bad = (out != r)
print("mismatches", int(bad.sum()), "of", M * N)
if bad.any():
    idx = bad.nonzero()[:12]
    for i, j in idx.tolist():
        print((i, j), "got", float(out[i, j]), "want", float(r[i, j]))
    print("rows with errors:", bad.any(1).nonzero().flatten()[:20].tolist(), "cols:", bad.any(0).nonzero().flatten()[:40].tolist())
torch.set_printoptions(linewidth=250, sci_mode=False)
d = (out - r)
for row in (0, 1, 2, 8, 9, 33):
    print("row", row, "out-r first 40:", d[row, :40].int().tolist())

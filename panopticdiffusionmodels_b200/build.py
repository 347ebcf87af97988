"""Build recipe for libpdm.so (hand-written CUDA for sm_100a + the C ABI of include/pdm.h).

    python -m panopticdiffusionmodels_b200.build

nvcc cross-compiles without a GPU.  The library is built IN-TREE
(panopticdiffusionmodels_b200/libpdm.so) so that it travels with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
TAG = os.environ.get("PDM_BUILD_TAG", "")  # development: a second library (e.g. a -DPDM_ATTN_TRACE build) next to the product one
OBJ = os.path.join(HERE, "build" + TAG)
LIB = os.path.join(HERE, f"libpdm{TAG}.so")
SOURCES = ["elementwise.cu", "gemm_simt.cu", "gemm_tc.cu", "attention_simt.cu", "attention_tc3.cu", "plan.cu", "vae.cu", "engine.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libpdm.so cannot be built")
    return nvcc


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    extra = os.environ.get("PDM_NVCC_EXTRA", "").split()  # development switches, e.g. -DPDM_ATTN_TRACE
    deps = [os.path.join(r, f) for r, _, fs in os.walk(CSRC) for f in fs] + [os.path.join(HERE, "..", "include", "pdm.h")]
    stamp = os.path.join(OBJ, "stamp.txt")
    digest = _digest(deps) + " ".join(extra)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()
    sources = list(SOURCES)
    if "-DPDM_ATTN_EXPERIMENTS" in extra:  # development: measured-and-rejected kernel variants kept for the record
        sources.append(os.path.join("experiments", "attention_tc4.cu"))

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src).replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    log = os.path.join(OBJ, "ptxas.log")
    with open(log, "w") as f:
        for obj, err in results:
            f.write(f"==== {os.path.basename(obj)}\n{err}\n")
            if verbose:
                print(err)
    cmd = [nvcc, "-shared", "-o", LIB, *[o for o, _ in results], "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

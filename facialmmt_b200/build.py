"""In-tree build of libfacialmmt_b200.so (hand-written sm_100a CUDA behind a C ABI).

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so that it travels with the
repository snapshot to the GPU box. No torch / libtorch linkage: the boundary is plain pointers (include/*.h).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = CSRC / "build"
LIB_PATH = PKG_DIR / "libfacialmmt_b200.so"

SOURCES = ["gemm.cu", "mlp_fused.cu", "mlp_stream.cu", "mlp_pair.cu", "kernels.cu", "attention.cu", "attn_fused.cu", "ln_qkv.cu", "ingest.cu",
           "umma_probe.cu", "engine.cu", "capi.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-DFMMT_BUILD",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-Xcompiler", "-ffp-contract=off",     # ingest.cu builds OpenCV's tap tables on the host: no FMA contraction there
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (needed to build libfacialmmt_b200.so)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link the shared library. Returns its path."""
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted((PKG_DIR.parent / "include").glob("*.h"))
    stamp = BUILD_DIR / "stamp.txt"
    digest = _digest(deps)
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    inc = ["-I", str(PKG_DIR.parent / "include"), "-I", str(CSRC)]
    procs = []
    objs = []
    for s in srcs:
        o = BUILD_DIR / (s.stem + ".o")
        objs.append(o)
        cmd = [nvcc, *NVCC_FLAGS, *inc, "-c", str(s), "-o", str(o)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s.name}:\n{out}")
        if verbose and out.strip():
            print(out, file=sys.stderr)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-o", str(LIB_PATH), *[str(o) for o in objs]]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    stamp.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

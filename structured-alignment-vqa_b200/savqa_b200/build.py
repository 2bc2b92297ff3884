"""In-tree build of libsavqa_b200.so (hand-written sm_100a CUDA behind a C ABI) with nvcc.

`python -m savqa_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU; the .so is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)  # structured-alignment-vqa_b200/
CSRC = os.path.join(ROOT, "csrc")
LIB_DIR = os.path.join(ROOT, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsavqa_b200.so")
SOURCES = ["api.cu", "elementwise.cu", "milnce.cu", "layernorm.cu", "gemm_tcgen05.cu", "gemm2_tcgen05.cu", "rowln_tcgen05.cu", "attn_simt.cu", "attn_tcgen05.cu", "attn_bwd_tcgen05.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: savqa_b200 needs the CUDA 12.9 toolkit to build its sm_100a kernels")


def _digest() -> str:
    h = hashlib.sha256()
    deps = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "savqa_b200.h")]
    for name in deps:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a into one shared library; returns its path.  Skips when up to date."""
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.stamp")
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose and out:
            print(out)
    if failed:
        raise RuntimeError("nvcc build of libsavqa_b200.so failed")
    link = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.run(link, check=True)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

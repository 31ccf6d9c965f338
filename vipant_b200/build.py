"""In-tree build of libvipant_b200.so (nvcc, sm_100a only).

    python -m vipant_b200.build [--force] [--verbose]

The shared library lands in ``vipant_b200/_lib/`` (git-ignored, shipped to the GPU box with the
snapshot).  nvcc cross-compiles without a GPU, so this also runs in the CPU-only container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libvipant_b200.so")
SOURCES = ["api.cu", "normalize.cu", "infonce_post.cu", "infonce_simt.cu", "infonce_tc.cu", "infonce_pair.cu", "retrieval.cu", "retrieval_fused.cu", "multilabel.cu", "encoder_tail.cu", "comm.cu", "p2p.cu"]
HEADERS = ["common.cuh", "simt_dot.cuh", "p2p.cuh", os.path.join("..", "..", "include", "vipant_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:  # one nvcc per translation unit, in parallel
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "--shared"] + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} (exit {p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static",
            "-Xcompiler", "-fPIC", "-o", LIB_PATH + ".tmp"] + objs + ["-ldl"]
    subprocess.run(link, check=True)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

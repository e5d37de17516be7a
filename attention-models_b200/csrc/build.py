"""Build libvq_b200.so in-tree with nvcc for sm_100a (no torch, no cmake).

    python attention-models_b200/csrc/build.py [--force]

The .so lands in attention-models_b200/lib/ (git-ignored, travels with gpurun snapshots).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libvq_b200.so")
SOURCES = ["vq_abi.cu", "vq_prep.cu", "vq_dist_simt.cu", "vq_dist_tc.cu", "vq_dist_tc16.cu", "vq_finish.cu", "vq_backward.cu", "vq_peer.cu", "vq_tokens.cu", "vq_prequant.cu"]
HEADERS = ["vq_common.cuh", "vq_kernels.h", "vq_tc_common.cuh", "vq_backward_body.cuh",
           os.path.join("..", "..", "include", "vq_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, instrument: bool = False, defines=(), suffix: str = "") -> str:
    """instrument=True builds lib/libvq_b200_instr.so with -DVQ_TC_INSTRUMENT (wait-cycle counters in the
    tensor-core kernel; diagnostic only, selected with VQ_B200_LIB=<path>)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    global LIB
    lib_path = LIB.replace(".so", "_instr.so") if instrument else LIB
    if suffix:      # experiment builds: extra -D flags, selected at run time with VQ_B200_LIB=<path>
        lib_path = LIB.replace(".so", f"_{suffix}.so")
    tag = "_instr" if instrument else (f"_{suffix}" if suffix else "")
    extra = [f"-D{d}" for d in defines]
    stamp = lib_path.replace(".so", ".stamp")
    digest = _digest() + "".join(extra)
    if not force and os.path.exists(lib_path) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return lib_path
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", tag + ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *(["-DVQ_TC_INSTRUMENT"] if instrument else []), *extra, "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0):
            print(f"--- {src}\n{out}", file=sys.stderr)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libvq_b200.so")
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path, *objs, "-cudart", "static"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout, file=sys.stderr)
        raise RuntimeError("link failed for libvq_b200.so")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return lib_path


if __name__ == "__main__":
    _defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    _suffix = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--suffix=")), "")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, instrument="--instrument" in sys.argv,
                defines=_defs, suffix=_suffix))

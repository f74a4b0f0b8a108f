"""In-tree nvcc build of libpxr.so for sm_100a (no JIT cache: the built .so
travels with the repo snapshot to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
REPO = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libpxr.so"
SOURCES = ["pxr_api.cu", "simt_kernels.cu", "score_tc.cu", "items_tc.cu", "sampling.cu", "novelty.cu", "diversity.cu"]
# score_tc.cu is compiled once per fusion_activation (pxr_act value -> -DPXR_TC_TU): object 0 = ReLU kernels (bf16 operands) +
# host side, objects 1-4 = the fused kernels of gelu / tanh / leaky_relu / silu, object 5 = the ReLU kernels for fp16 operands.
# (source, extra flags, object name); the longest units are started first
UNITS = [("score_tc.cu", [f"-DPXR_TC_TU={a}"], f"score_tc_act{a}.o") for a in (5, 1, 2, 3, 4)] + [(s, [], s + ".o") for s in SOURCES]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math" if False else "-DPXR_PRECISE_MATH", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
              "-I", str(REPO / "include"), "-I", str(CSRC)]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob("*.cuh")) + [REPO / "include" / "pxr.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objs = []
    bdir = PKG / "build"
    bdir.mkdir(exist_ok=True)
    procs = []
    for s, extra, oname in UNITS:
        o = bdir / oname
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((oname, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

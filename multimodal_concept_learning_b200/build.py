"""Build libmcl_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m multimodal_concept_learning_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcl_sm100.so")
SOURCES = ["api.cu", "scan_tc.cu", "scan_tc_m0.cu", "scan_tc_m1.cu", "scan_tc_m2.cu", "scan_tc_m3.cu", "gemm_tc.cu", "bwd_simt.cu", "scan_simt.cu", "merge.cu", "select.cu", "panel_scan.cu", "p2p_exchange.cu", "rowops.cu"]
HEADERS = ["common.cuh", "scan_tc_kernel.cuh", "rownorm.cuh", "rowstate.cuh", "kernels.h", "plan.h", "toplist.cuh", os.path.join("..", "..", "include", "mcl.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmcl_sm100.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, out: str = LIB, defines=()) -> str:
    """`out` / `defines` build experiment variants beside the product library (A/B timing).
    Every translation unit is compiled on its own thread (no relocatable device code is needed:
    kernels never call across files), then linked."""
    if out == LIB and not defines and not force and not is_stale():
        return LIB
    import concurrent.futures
    import tempfile
    nvcc = _nvcc()
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + [f"-D{d}" for d in defines]
    if verbose:
        cflags += ["-Xptxas", "-v"]
    with tempfile.TemporaryDirectory(prefix="mcl_build_") as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, src.replace(".cu", ".o"))
            cmd = [nvcc, *cflags, "-c", os.path.join(CSRC, src), "-o", obj]
            res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            return obj, cmd, res
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
            done = list(pool.map(compile_one, SOURCES))
        for obj, cmd, res in done:
            if verbose or res.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed with exit code {res.returncode}")
        link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
                "-Xcompiler", "-fPIC", "-o", out, *[d[0] for d in done], "-ldl"]
        res = subprocess.run(link, cwd=CSRC, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc link failed with exit code {res.returncode}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

"""Builds libeeseg_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m ee_semantic_segmentation_b200.build [--force] [--tuning]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the
gpurun snapshot. Nothing here depends on torch.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libeeseg_b200.so")
LIB_TUNING = os.path.join(HERE, "libeeseg_b200_tuning.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libeeseg_b200.so cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    files += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(f) for f in files)


def _compile(src, force, dep_m, log, tuning=False):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + (".tuning.o" if tuning else ".o"))
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), dep_m):
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DEESEG_TUNING"] if tuning else []) + ["-I", INCLUDE, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log[src] = r.stderr
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force=False, verbose=False, tuning=False):
    """tuning=True builds libeeseg_b200_tuning.so with -DEESEG_TUNING (in-kernel cycle counters and the
    include/eeseg_tuning.h hooks; load it with EESEG_LIB=...); the default product library has neither."""
    os.makedirs(OBJ, exist_ok=True)
    dep_m = _deps_mtime()
    log = {}
    LIB = LIB_TUNING if tuning else globals()["LIB"]
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, dep_m, log, tuning), sources()))
    if verbose:
        for k, v in log.items():
            print(k, "\n", v)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                        "-cudart", "static", "-lcuda" if False else "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tuning="--tuning" in sys.argv))

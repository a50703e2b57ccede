"""In-tree build of libcellseg_b200.so (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the CPU container and the built
.so travels to the GPU box with the repository snapshot.  Objects are cached by
source mtime under csrc/_build/.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "_build")
LIB = os.path.join(CSRC, "libcellseg_b200.so")

SOURCES = [
    "api_core.cu",
    "hsv_refine.cu",
    "unfold.cu",
    "select_topk.cu",
    "select_fast.cu",
    "select_reg.cu",
    "select_warp.cu",
    "paint.cu",
    "cc.cu",
    "fwd_fp32.cu",
    "fwd_tc.cu",
    "conv_gemm.cu",
    "stem_win.cu",
    "stem_ts.cu",
    "conv_halo.cu", "conv_ysum.cu", "conv_block.cu",
    "model.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libcellseg_b200.so")


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "cellseg_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def build(verbose=False, force=False):
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    hdr_m = _deps_mtime()
    objs, rebuilt = [], False
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(op)
        if (not force and os.path.exists(op)
                and os.path.getmtime(op) >= max(os.path.getmtime(sp), hdr_m)):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", op]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
        rebuilt = True
    if rebuilt or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libcellseg_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))

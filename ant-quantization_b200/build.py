"""Build libantq.so (hand-written CUDA for sm_100a + the C ABI of include/antq.h).

    python ant-quantization_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so is written next to the sources
(ant-quantization_b200/csrc/libantq.so) so that it travels with the tree; it
links against the CUDA runtime only -- no torch, no Python.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SUFFIX = os.environ.get("ANTQ_LIB_SUFFIX", "")
OUT = os.path.join(CSRC, "libantq%s.so" % SUFFIX)
SOURCES = ["antq_prepare.cu", "antq_stream.cu", "antq_pu.cu", "antq_short.cu", "antq_flat.cu", "antq_codes.cu", "antq_bwd.cu",
           "antq_calib.cu", "antq_gemm.cu", "antq_capi.cu"]
HEADERS = ["antq_common.cuh", os.path.join("..", "..", "include", "antq.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "--ftz=false", "--prec-div=true", "--prec-sqrt=true", "--fmad=false",
         "-Xptxas", "-v"] + os.environ.get("ANTQ_EXTRA_DEFS", "").split()


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(CSRC, src.replace(".cu", SUFFIX + ".o"))
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj


def build(force=False, verbose=False):
    if not (force or _stale()):
        return OUT
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = [NVCC, "-shared", "--cudart=static", "-o", OUT] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("built", OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)

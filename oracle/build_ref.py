"""Build oracle/_ref/ref_quant_cuda*.so from the reference's own CUDA sources (build container only).

    python oracle/build_ref.py

Needs /root/reference (mounted read-only in the build container); on the GPU box the prebuilt .so travels
with the repo snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).  No reference source is copied.
"""
import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
REF = "/root/reference/ant_quantization/quant"


def build(force=False):
    existing = glob.glob(os.path.join(OUT_DIR, "ref_quant_cuda*.so"))
    if existing and not force:
        return existing[0]
    if not os.path.exists(os.path.join(REF, "quant_kernel.cu")):
        raise FileNotFoundError(REF)
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    out = os.path.join(OUT_DIR, "ref_quant_cuda" + sysconfig.get_config_var("EXT_SUFFIX"))
    inc = []
    for p in ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]:
        inc += ["-I", p]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           '-DANTQ_REF_KERNEL_CU="%s/quant_kernel.cu"' % REF, '-DANTQ_REF_BINDING_CPP="%s/quant.cpp"' % REF,
           "-DTORCH_EXTENSION_NAME=ref_quant_cuda"] + inc + \
          [os.path.join(HERE, "ref_quant_cuda_wrap.cu"), "-o", out, "-L", libdir, "-lc10", "-ltorch", "-ltorch_cpu",
           "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-Xlinker", "-rpath", "-Xlinker", libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference kernel did not build:\n" + r.stderr[-3000:])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

"""Tuning builds: recompile one source (antq_stream.cu, or --src=<file>) with extra -D defines and link it with the
objects of the main build ->  csrc/libantq<suffix>.so  (select it with ANTQ_LIB_SUFFIX=<suffix>).

    python tools/build_variant.py _c15 -DANTQS_CONSUMERS=15 -DANTQS_STAGES=26
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ant-quantization_b200", "csrc")
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200"))
import build as B

suffix, defs = sys.argv[1], sys.argv[2:]
SRC = "antq_stream.cu"
if defs and defs[0].startswith("--src="):            # e.g. --src=antq_pu.cu
    SRC, defs = defs[0][6:], defs[1:]
obj = os.path.join(CSRC, SRC.replace(".cu", "%s.o" % suffix))
cmd = [B.NVCC] + B.FLAGS + defs + ["-c", os.path.join(CSRC, SRC), "-o", obj]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
objs = [os.path.join(CSRC, s.replace(".cu", ".o")) for s in B.SOURCES if s != SRC] + [obj]
out = os.path.join(CSRC, "libantq%s.so" % suffix)
subprocess.check_call([B.NVCC, "-shared", "--cudart=static", "-o", out] + objs)
print("built", out)

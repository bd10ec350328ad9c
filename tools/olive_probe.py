import sys, json
sys.path.insert(0, "/root/repo/tools")
import sweep
from antq import _lib
for fl, note in ((0, "pu-ovp"), (_lib.FLAG_NO_PU, "chain")):
    for sc in (1.0, 1.0 / 0.6028, 2.0):
        r = sweep.case(4096, "flint", 4, False, True, "row", "f16", flags=fl, alpha_scale=sc)
        print(note, sc, r["plan"], r["us"], r["frac"], flush=True)
for fl, note in ((_lib.FLAG_FORCE_PU, "pu-ovp signed"), (0, "chain signed")):
    for sc in (1.0, 1.5):
        r = sweep.case(4096, "flint", 4, True, True, "row", "f16", flags=fl, alpha_scale=sc)
        print(note, sc, r["plan"], r["us"], r["frac"], flush=True)

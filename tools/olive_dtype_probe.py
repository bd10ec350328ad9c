import sys, json
sys.path.insert(0, "/root/repo/tools")
import sweep
from antq import _lib
for dt in (sys.argv[1:] or ["f32", "bf16"]):
    for kind in ("flint", "int"):
        for signed in (True, False):
            for fl, nm in ((_lib.FLAG_FORCE_PU, "closed form"), (_lib.FLAG_NO_PU, "chain")):
                try:
                    r = sweep.case(4096, kind, 4, signed, True, "tensor", dt, flags=fl, alpha_scale=(1.0 if signed else 1.0 / 0.6028))
                    print(dt, kind, "s" if signed else "u", nm, r["plan"], r["us"], r["frac"], flush=True)
                except Exception as e:
                    print(dt, kind, signed, nm, "n/a", str(e)[:80], flush=True)

import sys, json
sys.path.insert(0, "/root/repo/tools")
import sweep
from antq import _lib
DT = sys.argv[1] if len(sys.argv) > 1 else "bf16"
for kind in ("flint", "int", "pot"):
    for gr in ("row", "tensor"):
        for fl, nm in ((0, "default"), (_lib.FLAG_FORCE_PU, "closed form"), (_lib.FLAG_FORCE_ROWS, "chain")):
            try:
                r = sweep.case(4096, kind, 4, True, False, gr, DT, flags=fl)
                print(kind, gr, nm, r["plan"], r["us"], r["frac"], flush=True)
            except Exception as e:
                print(kind, gr, nm, "n/a", str(e)[:60])

"""One-off probe of the closed-form kernels (4096 x 4096 fp16): short-row kernel build variants (ANTQ_LIB_SUFFIX) and, for
the stream kernel, per-row scales vs one scale vs per-row storage of one scale, with and without the exact redo
(ANTQ_DEBUG=4: measurement only, results are wrong).   python tools/pu_probe.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
sys.path.insert(0, os.path.join(%r, "ant-quantization_b200")); sys.path.insert(0, os.path.join(%r, "tools"))
import torch, antq
from antq import codebooks
from sweep import time_graph
dev = torch.device("cuda:0")
n = 4096
out = {}
for kind, bit, signed in (("int", 8, True), ("flint", 4, False), ("int", 6, True)):
    cb = antq.prepare_codebook(codebooks.ant_grid(kind, bit, signed).to(dev))
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [(torch.randn(n, n, device=dev, generator=g) * 0.02).to(torch.float16) for _ in range(6)]
    if not signed: xs = [x.abs() for x in xs]
    outs = [torch.empty_like(x) for x in xs]
    def run(view, alphas, per_row):
        def step():
            for i in range(len(xs)):
                antq.fakequant(view(xs[i]), alphas[i], cb, per_row, out=view(outs[i]))
        return round(time_graph(step, 20) / len(xs), 2)
    a_row = [(x.float().abs().amax(1) * 0.9).contiguous() for x in xs]
    a_one = [x.float().abs().max().reshape(1) * 0.9 for x in xs]
    a_same = [a.expand(n).contiguous() for a in a_one]
    key = "%%s%%d%%s" %% (kind, bit, "s" if signed else "u")
    out[key] = {"row": run(lambda t: t, a_row, True), "tensor": run(lambda t: t, a_one, False),
                "row_same_alpha": run(lambda t: t, a_same, True)}
    for gsz in (8, 16, 32, 64, 128, 256):
        a_g = [(x.view(-1, gsz).float().abs().amax(1) * 0.9).contiguous() for x in xs]
        out[key]["g%%d" %% gsz] = run(lambda t: t.view(-1, gsz), a_g, True)
    for gsz in (8, 32, 128):
        def stepd():
            for i in range(len(xs)):
                antq.fakequant_dynamic(xs[i].view(-1), cb, gsz, ratio=0.9, out=outs[i].view(-1))
        out[key]["dyn%%d" %% gsz] = round(time_graph(stepd, 20) / len(xs), 2)
print(json.dumps(out))
''' % (ROOT, ROOT)

for suffix, dbg in (("", "0"),):
    env = dict(os.environ, ANTQ_LIB_SUFFIX=suffix, ANTQ_DEBUG=dbg)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    print(json.dumps({"lib": suffix or "default", "debug": dbg, "us": json.loads(line[-1]) if line else r.stderr[-400:]}), flush=True)

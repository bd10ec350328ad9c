"""A small tour of the closed-form kernels for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py
Every path once on small shapes: long rows (near + wild + dead rows, queue overflow), SHORT mode, tile kernel with ragged
shapes, dynamic group scales, OliVe pairs (queue + dense list), the chain kernel, encode / decode, calibration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, antq
import antq_oracle as orc
from antq import _lib, codebooks
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)

def check(y, ref, what):
    a, b = y.cpu().numpy(), ref
    same = (a.view(np.uint16) == b.view(np.uint16)) | (np.isnan(a.astype(np.float32)) & np.isnan(b.astype(np.float32)))
    assert same.all(), (what, int((~same).sum()))
    print("ok", what, flush=True)

for kind, bit, signed in (("int", 8, True), ("flint", 4, False)):
    grid = orc.ant_grid(kind, bit, signed)
    cb = antq.prepare_codebook(torch.from_numpy(grid).to(dev))
    x = (rng.standard_normal((40, 4096)) * 0.02).astype(np.float32)
    x[rng.random(x.shape) < 0.01] *= 40
    x[3, 7], x[9] = np.nan, 0.0
    if not signed: x = np.abs(x)
    x = x.astype(np.float16)
    alpha = (np.abs(np.nan_to_num(x.astype(np.float32))).max(1) * 0.4).astype(np.float32)
    alpha[::5] = np.float32(0.05 * grid.max() / 8)          # tie rows: many near elements -> queue overflow
    alpha[11] = -1.0
    ref = orc.ant_forward(x, alpha, grid, per_row=True)
    xd, ad = torch.from_numpy(x).to(dev), torch.from_numpy(alpha).to(dev)
    check(antq.fakequant(xd, ad, cb, True, flags=_lib.FLAG_FORCE_PU), ref, "stream %s-%d" % (kind, bit))
    for g in (8, 24, 128, 504):
        if x.size % g: continue
        xs = x.reshape(-1, g); a = (np.abs(np.nan_to_num(xs.astype(np.float32))).max(1) * 0.7).astype(np.float32)
        check(antq.fakequant(torch.from_numpy(xs).to(dev), torch.from_numpy(a).to(dev), cb, True), orc.ant_forward(xs, a, grid, per_row=True), "rows of %d %s-%d" % (g, kind, bit))
    y, a_dyn = antq.fakequant_dynamic(xd.view(-1), cb, 32, ratio=0.9, return_alpha=True)
    check(y.view(-1, 32), orc.ant_forward(x.reshape(-1, 32), a_dyn.cpu().numpy(), grid, per_row=True), "dynamic group-32 %s-%d" % (kind, bit))
for signed in (True, False):
    g, o = orc.olive_grid("flint", 4, signed), orc.olive_outlier_grid(4, signed)
    cb = antq.prepare_codebook(torch.from_numpy(g).to(dev), torch.from_numpy(o).to(dev))
    x = (rng.standard_normal((16, 8192)) * 0.02).astype(np.float32)
    x[rng.random(x.shape) < 0.003] *= 12
    x[4:6][rng.random((2, 8192)) < 0.3] *= 20               # a third of the vectors hold an outlier: the dense pair list
    if not signed: x = np.abs(x)
    x = x.astype(np.float16)
    alpha = np.full(16, 0.06, dtype=np.float32)
    ref = orc.olive_forward(x, alpha, g, o, per_row=True)
    for fl in (0, _lib.FLAG_FORCE_PU):
        check(antq.fakequant(torch.from_numpy(x).to(dev), torch.from_numpy(alpha).to(dev), cb, True, ovp=True, flags=fl), ref, "olive signed=%s flags=%d" % (signed, fl))
torch.cuda.synchronize()
print("done")

"""Generate the golden fixtures in tests/golden/ by EXECUTING THE UNMODIFIED
REFERENCE PYTHON from /root/reference (build container only).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz, manifest.json

The reference's native half (a CUDA kernel) cannot run here; a literal
pure-torch restatement of its scan is injected as `quant_cuda`
(oracle/ref_harness.py::_scan_stub).  Everything above the kernel --
grid generators, _forward, OVP masking, search_mse, type selection,
quantize_model, autograd -- is the reference's own code, unmodified.

Fixtures are deliberately small (tensors of <= 24x40); they are committed so
the GPU box (which has no /root/reference) can check against them.
"""
import json
import os
import subprocess
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(4)


def npy(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------- grids ---
def gen_grids():
    out, meta = {}, []
    for tree, kinds in (("ant", ["int", "flint", "pot", "float1", "float2", "float3", "float4", "apot"]),
                        ("olive", ["int", "flint", "outlier"])):
        for kind in kinds:
            for signed in (True, False):
                for bit in range(3, 9):
                    key = "%s_%s_%s_%d" % (tree, kind, "s" if signed else "u", bit)
                    try:
                        if kind.startswith("float"):
                            g = rh.grid_of(tree, "float", bit, signed, int(kind[-1]))
                        else:
                            g = rh.grid_of(tree, kind, bit, signed)
                    except (AssertionError, TypeError) as e:
                        meta.append(dict(key=key, error=type(e).__name__))
                        continue
                    out[key] = npy(g.to(torch.float32))
                    meta.append(dict(key=key, n=int(g.numel())))
    return out, meta


# -------------------------------------------------------------- inputs ----
KAT_X = [-2, -1, -0.75, -0.4375, -0.3, -0.03125, 0, 0.03125, 0.07, 0.2, 0.4375, 0.45, 0.75, 0.8, 1, 1.3,
         float("nan"), float("inf"), -float("inf"), 3e4, -3e4, 10239.9, 10240.1, -10239.9, -10240.1,
         1e-30, -1e-30, 1e-42, 5.0, -5.0, 2.0000002, 123.456]


def weight_like(rows, cols, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(rows, cols, generator=g) * 0.02


def act_like(rows, cols, seed, relu=False, mult=20.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, cols, generator=g)
    idx = torch.randint(0, x.numel(), (max(1, x.numel() // 100),), generator=g)
    x.view(-1)[idx] *= mult
    return torch.relu(x) if relu else x


def tie_probe(grid, scale):
    """Inputs that land exactly on / next to every midpoint of the grid."""
    g = torch.unique(grid)
    mids = (g[:-1] + g[1:]) / 2
    pts = torch.cat([mids, torch.nextafter(mids, torch.tensor(1e9)), torch.nextafter(mids, torch.tensor(-1e9)),
                     g, g * 1.5, g * 2.5])
    return pts * scale


# ----------------------------------------------------- pinned forwards ----
def gen_forward_ant():
    out, meta = {}, []
    cases = []
    for kind, eb in (("int", None), ("flint", None), ("pot", None), ("float", 2), ("float", 3), ("apot", None)):
        for signed in (True, False):
            cases.append((kind, eb, 4, signed))
    for bit in (3, 5, 6, 8):
        cases.append(("int", None, bit, True))
        cases.append(("int", None, bit, False))
    for bit in (3, 5, 6):
        cases.append(("flint", None, bit, True))
        cases.append(("flint", None, bit, False))
    cases.append(("pot", None, 5, True))
    cases.append(("float", 3, 6, True))
    for ci, (kind, eb, bit, signed) in enumerate(cases):
        grid = rh.grid_of("ant", kind, bit, signed, eb)
        tag = "%s%s_%s_%d" % (kind, eb if eb else "", "s" if signed else "u", bit)
        # (a) per-row weights
        x = weight_like(24, 40, 100 + ci)
        if not signed:
            x = x.abs()
        x[3] = 0.0                                    # all-zero channel -> alpha 0 -> NaN row in the reference
        alpha = x.abs().max(1).values.unsqueeze(1) * 0.9
        alpha[5] = 1.0                                # a channel whose scale is far from its data
        q = rh.pin(rh.make_quantizer("ant", kind, bit, signed, is_input=False, rows=24), grid, alpha)
        y = q(x)
        out["row_%s_x" % tag] = npy(x); out["row_%s_alpha" % tag] = npy(alpha)
        out["row_%s_y" % tag] = npy(y); out["row_%s_grid" % tag] = npy(grid)
        # (b) per-tensor activations with outliers, KAT specials and tie probes
        xa = act_like(16, 40, 200 + ci, relu=not signed)
        a0 = torch.tensor(float(xa.abs().max()) * 0.8)
        s0 = a0 / grid.max()
        specials = torch.tensor(KAT_X, dtype=torch.float32)
        xa = torch.cat([xa.view(-1), specials, tie_probe(grid, s0)])
        q = rh.pin(rh.make_quantizer("ant", kind, bit, signed, is_input=True), grid, a0)
        ya = q(xa)
        out["ten_%s_x" % tag] = npy(xa); out["ten_%s_alpha" % tag] = npy(a0)
        out["ten_%s_y" % tag] = npy(ya); out["ten_%s_grid" % tag] = npy(grid)
        meta.append(dict(tag=tag, kind=kind, eb=eb, bit=bit, signed=signed))
    # the survey's KAT: alpha = 1.0, signed 4-bit, per-tensor
    for kind in ("int", "flint", "pot"):
        grid = rh.grid_of("ant", kind, 4, True)
        x = torch.tensor(KAT_X[:18], dtype=torch.float32)
        q = rh.pin(rh.make_quantizer("ant", kind, 4, True, is_input=True), grid, torch.tensor(1.0))
        out["kat_%s_x" % kind] = npy(x); out["kat_%s_y" % kind] = npy(q(x)); out["kat_%s_grid" % kind] = npy(grid)
    return out, meta


def gen_forward_olive():
    out, meta = {}, []
    ci = 0
    for kind in ("int", "flint"):
        for signed in (True, False):
            for bit in (4, 3, 6, 8):
                for no_outlier in (False, True):
                    if no_outlier and bit != 4:
                        continue
                    args = rh.default_args("olive", no_outlier=no_outlier)
                    grid = rh.grid_of("olive", kind, bit, signed)
                    outl = rh.grid_of("olive", "outlier", bit, signed).to(torch.float32)
                    tag = "%s_%s_%d%s" % (kind, "s" if signed else "u", bit, "_noout" if no_outlier else "")
                    ci += 1
                    # per-row weights with a few large entries
                    x = act_like(24, 40, 300 + ci, relu=not signed, mult=8.0) * 0.05
                    mean, std = x.mean(-1), x.std(-1)
                    alpha = torch.maximum((mean + 3 * std).abs(), (mean - 3 * std).abs()).unsqueeze(1)
                    q = rh.pin(rh.make_quantizer("olive", kind, bit, signed, is_input=False, rows=24, args=args),
                               grid, alpha, outl)
                    out["row_%s_x" % tag] = npy(x); out["row_%s_alpha" % tag] = npy(alpha)
                    out["row_%s_y" % tag] = npy(q(x)); out["row_%s_grid" % tag] = npy(grid)
                    out["row_%s_outliers" % tag] = npy(outl)
                    # per-tensor, ODD numel (torch.roll wrap-around), outlier pairs, specials
                    xa = act_like(15, 41, 400 + ci, relu=not signed, mult=30.0)
                    m, s = xa.mean(), xa.std()
                    a0 = torch.maximum((m + 3 * s).abs(), (m - 3 * s).abs())
                    s0 = a0 / grid.max()
                    pairs = torch.tensor([1, 100, 100, 1, 100, 200, 3, 5, 40, 0.5, -40, 7, 500, 0.1, 33, 9,
                                          float("nan"), 50, 60, float("inf"), 1e6, 2, -1e6, 70],
                                         dtype=torch.float32) * s0
                    if not signed:
                        pairs = pairs.abs()
                    xa = torch.cat([pairs, tie_probe(torch.cat([grid, outl]), s0), xa.view(-1)])
                    if xa.numel() % 2 == 0:
                        xa = torch.cat([xa, torch.tensor([100.0]) * s0])
                    q = rh.pin(rh.make_quantizer("olive", kind, bit, signed, is_input=True, args=args),
                               grid, a0, outl)
                    out["ten_%s_x" % tag] = npy(xa); out["ten_%s_alpha" % tag] = npy(a0)
                    out["ten_%s_y" % tag] = npy(q(xa)); out["ten_%s_grid" % tag] = npy(grid)
                    out["ten_%s_outliers" % tag] = npy(outl)
                    meta.append(dict(tag=tag, kind=kind, bit=bit, signed=signed, no_outlier=no_outlier))
    # survey KATs: flint-4 signed, scale 1 (alpha = 32)
    grid = rh.grid_of("olive", "flint", 4, True); outl = rh.grid_of("olive", "outlier", 4, True)
    for name, vals in (("even", [1, 100, 100, 1, 100, 200, 3, 5, 40, 0.5, -40, 7, 500, 0.1, 33, 9]),
                       ("odd", [100, 1, 2, 3, 4])):
        x = torch.tensor(vals, dtype=torch.float32)
        q = rh.pin(rh.make_quantizer("olive", "flint", 4, True, is_input=True), grid, torch.tensor(32.0), outl)
        out["kat_%s_x" % name] = npy(x); out["kat_%s_y" % name] = npy(q(x))
    out["kat_grid"] = npy(grid); out["kat_outliers"] = npy(outl)
    return out, meta


# ---------------------------------------------------------- calibration ---
def gen_calib():
    out, meta = {}, []
    ci = 0
    for tree, modes in (("ant", ["int", "flint", "ant-int-pot", "ant-int-pot-float", "ant-int-pot-flint",
                                 "ant-int-pot-float-flint"]),
                        ("olive", ["ant-int-flint", "int"])):
        for mode in modes:
            for is_input in (False, True):
                for bit in (4, 6, 8):
                    if bit != 4 and mode not in ("ant-int-pot-flint", "ant-int-flint"):
                        continue
                    ci += 1
                    args = rh.default_args(tree) if tree == "ant" else rh.default_args(tree, w_up=150, a_up=150)
                    if bit == 6 and tree == "ant":
                        args.w_low = args.a_low = 100            # quant_6bit_ptq.sh
                    if is_input:
                        x = act_like(16, 48, 500 + ci, relu=(ci % 2 == 0), mult=6.0)
                    else:
                        x = weight_like(16, 48, 500 + ci)
                        x[:, :4] *= 4.0
                    q = rh.make_quantizer(tree, mode, bit, is_signed=not is_input, is_input=is_input,
                                          args=args, rows=16 if not is_input else None, name="L%d" % ci)
                    y = q(x)
                    tag = "%s_%s_%s_%d" % (tree, mode, "in" if is_input else "w", bit)
                    out[tag + "_x"] = npy(x); out[tag + "_y"] = npy(y)
                    out[tag + "_alpha"] = npy(q.alpha.data); out[tag + "_grid"] = npy(q.quant_grid)
                    out[tag + "_mse"] = npy(q.mse)
                    if tree == "olive":
                        out[tag + "_outliers"] = npy(q.outliers.to(torch.float32))
                    meta.append(dict(tag=tag, tree=tree, mode=mode, is_input=is_input, bit=bit, chosen=q.mode,
                                     signed=bool(q.is_signed),
                                     low=args.a_low if is_input else args.w_low,
                                     up=args.a_up if is_input else args.w_up))
    return out, meta


# -------------------------------------------------------------- autograd --
def gen_autograd():
    out = {}
    for kind, is_input in (("flint", False), ("int", True), ("pot", False)):
        grid = rh.grid_of("ant", kind, 4, True)
        x = (weight_like(8, 32, 700) * 30).requires_grad_(True)
        alpha = (x.detach().abs().max(1).values.unsqueeze(1) * 0.8) if not is_input else torch.tensor(1.7)
        q = rh.pin(rh.make_quantizer("ant", kind, 4, True, is_input=is_input, rows=8), grid, alpha)
        y = q(x)
        g = weight_like(8, 32, 701) * 10
        (y * g).sum().backward()
        tag = "%s_%s" % (kind, "in" if is_input else "w")
        out[tag + "_x"] = npy(x); out[tag + "_alpha"] = npy(alpha); out[tag + "_grid"] = npy(grid)
        out[tag + "_gout"] = npy(g); out[tag + "_y"] = npy(y)
        out[tag + "_gx"] = npy(x.grad); out[tag + "_galpha"] = npy(q.alpha.grad)
    return out


# ----------------------------------------------------------- model level --
MODEL_SCRIPT = r'''
import sys, os, json, types
import numpy as np, torch, torch.nn as nn
sys.path.insert(0, %(oracle)r)
import ref_harness as rh
rh.install_stub()
tree = %(tree)r
sys.path.append(rh.TREES[tree])
if tree == "ant":
    rh.ensure_gloo_group()
from quant_model import *
from quant_utils import *
torch.manual_seed(0)
class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.features = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1, bias=False), nn.ReLU())
        self.pool = nn.AdaptiveAvgPool2d(2)
        self.blocks = nn.ModuleList([nn.Linear(32, 32), nn.Linear(32, 32)])
        self.head = nn.Linear(32, 10)
    def forward(self, x):
        x = self.pool(self.features(x)).flatten(1)
        for b in self.blocks:
            x = torch.relu(b(x))
        return self.head(x)
net = Net().eval()
x = torch.randn(4, 3, 8, 8)
args = types.SimpleNamespace(mode=%(mode)r, wbit=4, abit=4, w_up=150, a_up=150, w_low=75, a_low=75, percent=100, search=False, no_outlier=False)
set_quantizer(args)
qnet = quantize_model(net)
enable_quantization(qnet)
with torch.no_grad():
    y_cal = qnet(x)          # first forward calibrates
    y = qnet(x)
sd = qnet.state_dict()
out = {"x": x.numpy(), "y_cal": y_cal.numpy(), "y": y.numpy(), "y_fp32": net(x).detach().numpy()}
for k, v in net.state_dict().items():
    out["fp32/" + k] = v.numpy()
for k, v in sd.items():
    out["sd/" + k] = v.detach().to(torch.float32).numpy()
modes = {n: m.mode for n, m in qnet.named_modules() if isinstance(m, TensorQuantizer)}
signed = {n: bool(m.is_signed) for n, m in qnet.named_modules() if isinstance(m, TensorQuantizer)}
np.savez_compressed(%(dst)r, **out)
json.dump({"modes": modes, "signed": signed, "keys": list(sd.keys())}, open(%(dst_json)r, "w"), indent=1)
'''


def gen_models():
    for tree, mode in (("ant", "ant-int-pot-flint"), ("olive", "ant-int-flint")):
        dst = os.path.join(HERE, "model_%s.npz" % tree)
        script = MODEL_SCRIPT % dict(oracle=os.path.join(ROOT, "oracle"), tree=tree, mode=mode, dst=dst,
                                     dst_json=os.path.join(HERE, "model_%s.json" % tree))
        subprocess.check_call([sys.executable, "-c", script], cwd=HERE)


def main():
    manifest = {}
    for name, fn in (("grids", gen_grids), ("forward_ant", gen_forward_ant),
                     ("forward_olive", gen_forward_olive), ("calib", gen_calib)):
        data, meta = fn()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        manifest[name] = meta
        print(name, len(data), "arrays")
    np.savez_compressed(os.path.join(HERE, "autograd.npz"), **gen_autograd())
    gen_models()
    manifest["reference_commit"] = "bc840673b614fac644081f3169a9c81dff2d8dc1"
    manifest["torch"] = torch.__version__
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

"""Kernel-only timing of antq.fakequant on rotating buffers (CUDA graph, CUDA events)."""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200"))
import torch, antq
from antq import codebooks, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4096); ap.add_argument("--cols", type=int, default=4096)
ap.add_argument("--dtype", default="f16"); ap.add_argument("--kind", default="flint"); ap.add_argument("--bit", type=int, default=4)
ap.add_argument("--unsigned", action="store_true"); ap.add_argument("--per-tensor", action="store_true")
ap.add_argument("--olive", action="store_true"); ap.add_argument("--flat", action="store_true")
ap.add_argument("--nb", type=int, default=8); ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--torch-copy", action="store_true"); ap.add_argument("--torch-read", action="store_true")
ap.add_argument("--alpha-mult", type=float, default=0.9); ap.add_argument("--tag", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
dt = {"f16": torch.float16, "f32": torch.float32, "bf16": torch.bfloat16}[a.dtype]
signed = not a.unsigned
if a.olive:
    cb = antq.prepare_codebook(codebooks.olive_grid(a.kind, a.bit, signed).to(dev), codebooks.olive_outliers(a.bit, signed).to(dev))
else:
    cb = antq.prepare_codebook(codebooks.ant_grid(a.kind, a.bit, signed).to(dev))
g = torch.Generator(device="cpu").manual_seed(0)
xs, als, outs = [], [], []
for i in range(a.nb):
    x = (torch.randn(a.rows, a.cols, generator=g) * 0.02)
    if not signed: x = x.abs()
    x = x.to(dt).to(dev)
    al = (x.float().abs().amax(1) * a.alpha_mult) if not a.per_tensor else (x.float().abs().max().reshape(1) * a.alpha_mult)
    if a.olive: al = al * 0 + (3 * x.float().std())
    xs.append(x); als.append(al.contiguous()); outs.append(torch.empty_like(x))
flags = _lib.FLAG_FORCE_FLAT if a.flat else 0
rd = [torch.empty(a.rows, dtype=dt, device=dev) for _ in range(a.nb)]
def step():
    if a.torch_copy:
        for i in range(a.nb): outs[i].copy_(xs[i])
        return
    if a.torch_read:
        for i in range(a.nb): torch.amax(xs[i], dim=1, out=rd[i])
        return
    for i in range(a.nb):
        antq.fakequant(xs[i], als[i], cb, not a.per_tensor, ovp=a.olive, out=outs[i], flags=flags)
step(); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr): step()
for _ in range(3): gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps): gr.replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (a.reps * a.nb)
nbytes = a.rows * a.cols * xs[0].element_size() * 2
print(json.dumps({"tag": a.tag, "plan": antq.fakequant_plan(xs[0], cb, not a.per_tensor, ovp=a.olive, flags=flags),
                  "shape": [a.rows, a.cols], "dtype": a.dtype, "kind": a.kind, "bit": a.bit, "signed": signed,
                  "per_tensor": a.per_tensor, "olive": a.olive, "us_per_launch": round(us, 2),
                  "GBps": round(nbytes / us / 1e3, 1), "frac_of_6548.8": round(nbytes / us / 1e3 / 6548.8, 3),
                  "warps_env": os.environ.get("ANTQ_ROWS_WARPS", "")}))

"""The 'kernel to beat' on the same B200: the reference's own op sequence (7 elementwise passes around its
scan kernel, A/antquant/quant_modules.py:535-551) with its own CUDA kernel (oracle/_ref), timed with CUDA
events on the headline tensor, fp32 (the reference kernel has no fp16 dispatch).  Reported in profiles/."""
import glob, importlib.util, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, antq
from antq import codebooks
so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_quant_cuda*.so"))
spec = importlib.util.spec_from_file_location("ref_quant_cuda", so[0]); refk = importlib.util.module_from_spec(spec); spec.loader.exec_module(refk)
dev = torch.device("cuda:0"); N = 4096; NB = 8
g = codebooks.ant_grid("flint", 4, True).to(dev)
xs = [(torch.randn(N, N, device=dev) * 0.02) for _ in range(NB)]
als = [(x.abs().amax(1) * 0.9).unsqueeze(1) for x in xs]
def ref_step():
    for x, a in zip(xs, als):
        scale = a / torch.max(g)
        d = (x.view(N, -1) / scale).view(x.shape)
        q, _ = refk.quant(d.view(-1), g); q = q.view(x.shape)
        t = (q - d) + d
        y = (t.view(N, -1) * scale).view(x.shape)
cb = antq.prepare_codebook(g); outs = [torch.empty_like(x) for x in xs]
def our_step():
    for x, a, o in zip(xs, als, outs):
        antq.fakequant(x, a, cb, True, out=o)
res = {}
for name, fn in (("reference_sequence_fp32", ref_step), ("antq_fp32", our_step)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (10 * NB)
    res[name] = {"us_per_tensor": round(us, 1), "algorithmic_GBps": round(N * N * 8 / us / 1e3, 1)}
print(json.dumps(res))

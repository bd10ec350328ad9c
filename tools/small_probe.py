"""Per-tensor fake-quant of SMALL tensors (BERT-base activations are 3.1 M elements): the persistent TMA-staged kernels
against the closed-form tile kernel (ANTQ_FLAG_FORCE_TILE), inside a CUDA graph.   python tools/small_probe.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, antq
from antq import _lib, codebooks
from sweep import time_graph
dev = torch.device("cuda:0")
for kind, bit, signed in (("flint", 4, True), ("int", 4, True), ("int", 8, True), ("flint", 4, False)):
    cb = antq.prepare_codebook(codebooks.ant_grid(kind, bit, signed).to(dev))
    for numel in (1 << 20, 32 * 128 * 768, 1 << 22, 1 << 23, 32 * 128 * 3072, 1 << 24, 1 << 25):
        nb = max(2, min(64, (300 << 20) // (numel * 4)))
        g = torch.Generator(device="cuda").manual_seed(1)
        xs = [(torch.randn(numel, device=dev, generator=g) * 0.02).to(torch.float16) for _ in range(nb)]
        if not signed: xs = [x.abs() for x in xs]
        outs = [torch.empty_like(x) for x in xs]
        al = [x.float().abs().max().reshape(1) * 0.9 for x in xs]
        res = {}
        for name, fl in (("default", 0), ("tile", _lib.FLAG_FORCE_TILE)):
            def step():
                for i in range(nb):
                    antq.fakequant(xs[i], al[i], cb, False, out=outs[i], flags=fl)
            res[name] = round(time_graph(step, 20) / nb, 2)
            res[name + "_plan"] = antq.fakequant_plan(xs[0], cb, False, flags=fl)
        assert torch.equal(antq.fakequant(xs[0], al[0], cb, False), antq.fakequant(xs[0], al[0], cb, False, flags=_lib.FLAG_FORCE_TILE))
        print(json.dumps({"grid": "%s-%d-%s" % (kind, bit, "s" if signed else "u"), "numel": numel, "roofline_us": round(numel * 4 / 6538.9e3, 2), **res}), flush=True)

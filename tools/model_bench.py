"""Model-shaped throughput of the fake-quant path THROUGH THE LAYER API (quantize_model + the *Quantizer wrappers), on
random-init architectures with synthetic inputs (BASELINE.json C2 / C4; reference hooks A/ImageNet/main.py:117-128,
O/llm/run_clm.py:603-613):

    python tools/model_bench.py opt      # one OPT-6.7B decoder layer (hidden 4096, ffn 16384, 32 heads), seq 2048, OliVe 4-bit
    python tools/model_bench.py resnet   # torchvision ResNet-50, batch 256 x 3 x 224 x 224 fp16, ANT flint-4 W + A
    python tools/model_bench.py bert     # BERT-base-shaped encoder (12 x 768, ffn 3072), batch 32 x seq 128, ANT int/flint-4 W + A

Each prints one JSON object: forward time of the plain fp16 model, of the quantized model (weights re-quantized every
forward like the reference / eval-mode weight cache / fused tcgen05 Linear), the algorithmic bytes the fake-quant path
touches per forward (sizeof(in) + sizeof(out) per quantized element) and the in-situ rate
bytes / (t_quantized - t_plain) against the measured HBM peak.  Eager PyTorch, CUDA events: Python and launch overheads of
the layer API are inside the numbers; the `cuda_graph` entries replay the same forwards (plain and quantized, weight cache
on) captured in a CUDA graph, which is how small-tensor models (BERT-base: 3 M-element activations) should be served.
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
which = sys.argv[1] if len(sys.argv) > 1 else "opt"
tree = "olive" if which == "opt" else "ant"
sys.path.append(os.path.join(ROOT, "ant-quantization_b200", tree, "antquant"))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from quant_model import *  # noqa: E402,F401,F403
from quant_utils import *  # noqa: E402,F401,F403
import antq.layers as L  # noqa: E402

dev = torch.device("cuda:0")
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps            # ms


class OPTLayer(nn.Module):
    """OPT decoder layer (pre-LN, ReLU FFN): the six nn.Linear the OliVe scripts quantize in every layer."""

    def __init__(self, h=4096, ffn=16384, heads=32):
        super().__init__()
        self.heads = heads
        self.ln1, self.ln2 = nn.LayerNorm(h), nn.LayerNorm(h)
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
        self.fc1, self.fc2 = nn.Linear(h, ffn), nn.Linear(ffn, h)

    def forward(self, x):
        B, S, H = x.shape
        y = self.ln1(x)
        sp = lambda t: t.view(B, S, self.heads, H // self.heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(sp(self.q_proj(y)), sp(self.k_proj(y)), sp(self.v_proj(y)), is_causal=True)
        x = x + self.out_proj(a.transpose(1, 2).reshape(B, S, H))
        return x + self.fc2(F.relu(self.fc1(self.ln2(x))))


class BertLayer(nn.Module):
    """BERT-base encoder layer: the six nn.Linear the ANT BERT script quantizes (A/BERT/run_glue.py:538-546)."""

    def __init__(self, h=768, ffn=3072, heads=12):
        super().__init__()
        self.heads = heads
        self.q, self.k, self.v, self.o = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
        self.ln1, self.ln2 = nn.LayerNorm(h), nn.LayerNorm(h)
        self.fc1, self.fc2 = nn.Linear(h, ffn), nn.Linear(ffn, h)

    def forward(self, x):
        B, S, H = x.shape
        sp = lambda t: t.view(B, S, self.heads, H // self.heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(sp(self.q(x)), sp(self.k(x)), sp(self.v(x)))
        x = self.ln1(x + self.o(a.transpose(1, 2).reshape(B, S, H)))
        return self.ln2(x + self.fc2(F.gelu(self.fc1(x))))


def graphed(fn):
    """fn() captured in a CUDA graph (after a warm-up on a side stream); returns the replay callable."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


class quant_bytes:
    """Algorithmic bytes of one forward = sizeof(in) + sizeof(out) of every fake-quant LAUNCH (identical activation
    quantizers fed the same tensor share one: antq/quantizer.py), counted by wrapping the op for one pass."""

    def __init__(self, qmodel, per_forward_weights):
        from antq import ops
        self.ops, self.real, self.tot = ops, ops.fakequant, [0]

        def counting(x, *a, **k):
            self.tot[0] += 2 * x.numel() * x.element_size()
            return self.real(x, *a, **k)
        ops.fakequant = counting

    def done(self):
        self.ops.fakequant = self.real
        return self.tot[0]


def run(model, x, args, fused_ok, graph=True):
    res = {}
    with torch.no_grad():
        res["plain_ms"] = timeit(lambda: model(x))
        if graph:
            res["cuda_graph"] = {"plain_ms": round(timeit(graphed(lambda: model(x))), 3)}
        set_quantizer(args)  # noqa: F405
        q = quantize_model(model).to(dev).eval()  # noqa: F405
        enable_quantization(q)  # noqa: F405
        q(x)                                                           # calibration forward
        torch.cuda.synchronize()
        for name, cache, fused in (("requant_every_forward", False, False), ("weight_cache", True, False)) + \
                ((("fused_tcgen05_linear", True, True),) if fused_ok else ()):
            L.CACHE_WEIGHTS, L.FUSED_LINEAR = cache, fused
            for m in q.modules():
                if hasattr(m, "invalidate_weight_cache"):
                    m.invalidate_weight_cache()
            q(x)
            qb = quant_bytes(q, not cache)
            q(x)
            tot = qb.done()
            ms = timeit(lambda: q(x))
            extra = ms - res["plain_ms"]
            res[name] = {"forward_ms": round(ms, 3), "quant_bytes_per_forward": tot,
                         "in_situ_GBps": round(tot / (extra * 1e-3) / 1e9, 1) if extra > 0 else None,
                         "frac_of_hbm_peak": round(tot / (extra * 1e-3) / 1e9 / PEAK, 3) if extra > 0 else None}
        L.CACHE_WEIGHTS, L.FUSED_LINEAR = True, False
        if graph:
            for m in q.modules():
                if hasattr(m, "invalidate_weight_cache"):
                    m.invalidate_weight_cache()
            q(x)
            qb = quant_bytes(q, False)
            q(x)
            tot = qb.done()
            ms = timeit(graphed(lambda: q(x)))
            extra = ms - res["cuda_graph"]["plain_ms"]
            res["cuda_graph"].update({"weight_cache_ms": round(ms, 3), "quant_bytes_per_forward": tot,
                                      "in_situ_GBps": round(tot / (extra * 1e-3) / 1e9, 1) if extra > 0 else None,
                                      "frac_of_hbm_peak": round(tot / (extra * 1e-3) / 1e9 / PEAK, 3) if extra > 0 else None})
    res["plain_ms"] = round(res["plain_ms"], 3)
    return res


torch.manual_seed(0)
if which == "opt":
    model = OPTLayer().to(dev).half().eval()
    x = torch.randn(1, 2048, 4096, device=dev, dtype=torch.float16)
    args = types.SimpleNamespace(mode="ant-int-flint", wbit=4, abit=4, w_up=250, a_up=250, w_low=75, a_low=75, percent=100,
                                 search=False, no_outlier=False)
    out = {"workload": "OPT-6.7B decoder layer (h 4096, ffn 16384, 32 heads), seq 2048, batch 1, fp16, OliVe 4-bit W + A (outlier-victim pairs)"}
    out.update(run(model, x, args, fused_ok=False))
    # the same layer with ANT-style grids (no outlier pairs) can take the fused tcgen05 Linear
    args2 = types.SimpleNamespace(**{**vars(args), "no_outlier": True})
    out["no_outlier_variant"] = run(model, x, args2, fused_ok=True)
elif which == "bert":
    model = nn.Sequential(*[BertLayer() for _ in range(12)]).to(dev).half().eval()
    x = torch.randn(32, 128, 768, device=dev, dtype=torch.float16)
    args = types.SimpleNamespace(mode="ant-int-flint", wbit=4, abit=4, w_up=150, a_up=150, w_low=75, a_low=75, percent=100, search=False)
    out = {"workload": "BERT-base-shaped encoder (12 layers x 768, ffn 3072, 12 heads), batch 32 x seq 128, fp16, ANT int/flint-4 W + A"}
    out.update(run(model, x, args, fused_ok=False))
else:
    import torchvision
    model = torchvision.models.resnet50(weights=None).to(dev).half().eval()
    x = torch.randn(256, 3, 224, 224, device=dev, dtype=torch.float16)
    args = types.SimpleNamespace(mode="flint", wbit=4, abit=4, w_up=150, a_up=150, w_low=75, a_low=75, percent=100, search=False)
    out = {"workload": "torchvision ResNet-50, batch 256 x 3 x 224 x 224, fp16, ANT flint-4 W + A (per-channel weights, per-tensor activations)"}
    out.update(run(model, x, args, fused_ok=False))
out["hbm_peak_GBps"] = PEAK
print(json.dumps(out))

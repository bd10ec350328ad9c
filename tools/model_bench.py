"""Model-shaped throughput of the fake-quant path THROUGH THE LAYER API (quantize_model + the *Quantizer wrappers), on
random-init architectures with synthetic inputs (BASELINE.json C2 / C4; reference hooks A/ImageNet/main.py:117-128,
O/llm/run_clm.py:603-613):

    python tools/model_bench.py opt      # one OPT-6.7B decoder layer (hidden 4096, ffn 16384, 32 heads), seq 2048, OliVe 4-bit
    python tools/model_bench.py resnet   # torchvision ResNet-50, batch 256 x 3 x 224 x 224 fp16, ANT flint-4 W + A

Each prints one JSON object: forward time of the plain fp16 model, of the quantized model (weights re-quantized every
forward like the reference / eval-mode weight cache / fused tcgen05 Linear), the algorithmic bytes the fake-quant path
touches per forward (sizeof(in) + sizeof(out) per quantized element) and the in-situ rate
bytes / (t_quantized - t_plain) against the measured HBM peak.  Eager PyTorch, CUDA events, no CUDA graph: Python and
launch overheads of the layer API are inside the numbers.
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


def quant_bytes(qmodel, per_forward_weights):
    """Algorithmic bytes of one forward: every activation quantizer call, plus the weights if they are re-quantized."""
    tot = [0]
    hs = []
    for m in qmodel.modules():
        if isinstance(m, TensorQuantizer):  # noqa: F405
            if m.is_input or per_forward_weights:
                hs.append(m.register_forward_hook(lambda mod, inp, out: tot.__setitem__(0, tot[0] + 2 * inp[0].numel() * inp[0].element_size())))
    return tot, hs


def run(model, x, args, fused_ok):
    res = {}
    with torch.no_grad():
        res["plain_ms"] = timeit(lambda: model(x))
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
            tot, hs = quant_bytes(q, not cache)
            q(x)
            for h in hs:
                h.remove()
            ms = timeit(lambda: q(x))
            extra = ms - res["plain_ms"]
            res[name] = {"forward_ms": round(ms, 3), "quant_bytes_per_forward": tot[0],
                         "in_situ_GBps": round(tot[0] / (extra * 1e-3) / 1e9, 1) if extra > 0 else None,
                         "frac_of_hbm_peak": round(tot[0] / (extra * 1e-3) / 1e9 / PEAK, 3) if extra > 0 else None}
        L.CACHE_WEIGHTS, L.FUSED_LINEAR = True, False
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
else:
    import torchvision
    model = torchvision.models.resnet50(weights=None).to(dev).half().eval()
    x = torch.randn(256, 3, 224, 224, device=dev, dtype=torch.float16)
    args = types.SimpleNamespace(mode="flint", wbit=4, abit=4, w_up=150, a_up=150, w_low=75, a_low=75, percent=100, search=False)
    out = {"workload": "torchvision ResNet-50, batch 256 x 3 x 224 x 224, fp16, ANT flint-4 W + A (per-channel weights, per-tensor activations)"}
    out.update(run(model, x, args, fused_ok=False))
out["hbm_peak_GBps"] = PEAK
print(json.dumps(out))

"""Wall-clock of one quantizer initialisation (type search + alpha search = the reference's _init_quant_para,
A/antquant/quant_modules.py:468-533) on B200, and of the steady-state forward that follows."""
import os, sys, time, types, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.append(os.path.join(ROOT, "ant-quantization_b200", "ant", "antquant"))
import torch
from quant_modules import TensorQuantizer

dev = torch.device("cuda:0")
args = types.SimpleNamespace(mode="ant-int-pot-flint", wbit=4, abit=4, w_up=150, a_up=150, w_low=75, a_low=75,
                             percent=100, search=False, no_outlier=False)
out = []
for name, shape, is_input, dt in (("weight 4096x4096 fp16 per-channel", (4096, 4096), False, torch.float16),
                                  ("weight 4096x4096 fp32 per-channel", (4096, 4096), False, torch.float32),
                                  ("activation 32x2048x4096 fp16 per-tensor", (32 * 2048, 4096), True, torch.float16)):
    torch.manual_seed(0)
    x = (torch.randn(*shape, device=dev) * 0.02).to(dt)
    if is_input:
        x = x.abs()
    best = None
    for attempt in range(4):                       # the fastest warm initialisation is the one reported
        q = TensorQuantizer(mode=args.mode, bit=4, is_signed=not is_input, is_enable=True, is_input=is_input, args=args).to(dev)
        if not is_input:
            q.alpha.data = torch.ones([shape[0], 1], device=dev)
        q.enable_quantization(name)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = q(x)
        torch.cuda.synchronize(); t_init = time.perf_counter() - t0
        if attempt > 0:
            best = t_init if best is None else min(best, t_init)
    t_init = best
    for _ in range(3): q(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): q(x)
    torch.cuda.synchronize(); t_fwd = (time.perf_counter() - t0) / 200
    out.append({"tensor": name, "chosen": q.mode, "init_ms": round(t_init * 1e3, 2), "steady_forward_us": round(t_fwd * 1e6, 1),
                "elements": x.numel()})
# per-call host overhead of the module API: a tensor so small that the kernel is negligible, 2000 back-to-back calls
x = (torch.randn(64, 512, device=dev) * 0.02).to(torch.float16)
q = TensorQuantizer(mode="flint", bit=4, is_signed=True, is_enable=True, is_input=False, args=args).to(dev)
q.alpha.data = torch.ones([64, 1], device=dev)
q.enable_quantization("tiny")
with torch.no_grad():
    q(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2000): q(x)
    torch.cuda.synchronize(); t_call = (time.perf_counter() - t0) / 2000
out.append({"tensor": "64x512 fp16 per-channel (host overhead probe)", "steady_forward_us": round(t_call * 1e6, 2)})
print(json.dumps(out))

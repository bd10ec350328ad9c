import sys, os, time, types
ROOT = "/root/repo"
sys.path.append(os.path.join(ROOT, "ant-quantization_b200", "ant", "antquant"))
import torch
from quant_model import *
from quant_utils import *
import antq.quantizer as Q
dev = torch.device("cuda:0")
args = types.SimpleNamespace(mode="flint", wbit=4, abit=4, w_up=150, a_up=150, w_low=75, a_low=75, percent=100, search=False)
qs = []
x = torch.randn(64, 512, device=dev).half()
for i in range(3):
    q = TensorQuantizer(mode="flint", bit=4, is_signed=True, is_enable=True, is_input=True, args=args).to(dev)
    q.enable_quantization("a%d" % i)
    with torch.no_grad():
        q(x)
    qs.append(q)
for share in (False, True, False, True):
    Q.SHARE_INPUT_QUANT = share
    with torch.no_grad():
        for _ in range(200):
            for q in qs: q(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 3000
        for _ in range(n):
            for q in qs: q(x)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / (3 * n) * 1e6
    print("share", share, "us per quantizer call %.2f" % dt, flush=True)

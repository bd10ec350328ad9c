"""A few launches of antq_linear_p4 at the OPT shape (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200"))
import torch, antq
from antq import codebooks
dev = torch.device("cuda:0")
cb = antq.prepare_codebook(codebooks.ant_grid("flint", 4, True).to(dev))
M, N, K = 2048, 4096, 4096
w = (torch.randn(N, K, device=dev) * 0.02).half()
al = (w.float().abs().amax(1) * 0.9).contiguous()
codes, _ = antq.encode_p4(w, al, cb, True)
x = torch.randn(M, K, device=dev).half()
for _ in range(4):
    y = antq.linear_p4(x, codes, al, cb, N)
torch.cuda.synchronize()

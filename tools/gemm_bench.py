"""antq_linear_p4 (dequant-fused tcgen05 GEMM) vs cuBLAS F.linear on the already fake-quantized fp16 weight, and vs
decode + F.linear.  CUDA events, rotating operands, TFLOP/s against MEASURED_PEAKS.json's bf16 GEMM peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200"))
import torch, torch.nn.functional as F
import antq
from antq import codebooks

dev = torch.device("cuda:0")
try:
    P = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peak, peak_s = P["bf16_tflops"], P.get("bf16_tflops_sustained")
except Exception:
    peak, peak_s = 1590.0, 1400.0


def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


out = []
cb = antq.prepare_codebook(codebooks.ant_grid("flint", 4, True).to(dev))
for dt in (torch.float16, torch.bfloat16):
    for M, N, K in ((2048, 4096, 4096), (2048, 16384, 4096), (2048, 4096, 16384), (8192, 4096, 4096), (128, 4096, 4096)):
        nb = 4
        ws = [(torch.randn(N, K, device=dev) * 0.02).to(dt) for _ in range(nb)]
        als = [(w.float().abs().amax(1) * 0.9).contiguous() for w in ws]
        wqs = [antq.fakequant(w, a, cb, True) for w, a in zip(ws, als)]
        cds = [antq.encode_p4(w, a, cb, True)[0] for w, a in zip(ws, als)]
        xs = [torch.randn(M, K, device=dev).to(dt) for _ in range(nb)]
        i = [0]
        def fused():
            k = i[0] = (i[0] + 1) % nb
            return antq.linear_p4(xs[k], cds[k], als[k], cb, N)
        def cublas():
            k = i[0] = (i[0] + 1) % nb
            return F.linear(xs[k], wqs[k])
        def decode_then():
            k = i[0] = (i[0] + 1) % nb
            return F.linear(xs[k], antq.decode_p4(cds[k], als[k], cb, (N, K), dt, True))
        def requant_then():                       # what the reference does every forward: re-fake-quantize, then GEMM
            k = i[0] = (i[0] + 1) % nb
            return F.linear(xs[k], antq.fakequant(ws[k], als[k], cb, True))
        xcb = antq.prepare_codebook(codebooks.ant_grid("flint", 4, False).to(dev))
        xas = [x.float().abs().max().reshape(1) * 0.8 for x in xs]
        xqs = [antq.fakequant(x.abs(), a, xcb, False) for x, a in zip(xs, xas)]
        def fused_fp8():                          # W4A4 as e4m3 levels (includes the level-conversion kernel)
            k = i[0] = (i[0] + 1) % nb
            return antq.linear_p4_fp8(xqs[k], xas[k], xcb, cds[k], als[k], cb, N)
        fl = 2.0 * M * N * K
        r = {"M": M, "N": N, "K": K, "dtype": str(dt).split(".")[1]}
        for name, fn in (("fused_tcgen05_fp8_w4a4", fused_fp8), ("fused_tcgen05", fused), ("cublas_on_fp16_weight", cublas), ("decode_p4+cublas", decode_then),
                         ("fakequant+cublas", requant_then)):
            us = timeit(fn, 20)
            r[name] = {"us": round(us, 1), "tflops": round(fl / us / 1e6, 1), "frac_of_measured_bf16_peak": round(fl / us / 1e6 / peak, 3)}
        ref = F.linear(xs[0], wqs[0]).float()
        err = (antq.linear_p4(xs[0], cds[0], als[0], cb, N).float() - ref).norm() / ref.norm()
        r["rel_diff_vs_cublas"] = float(err)
        r["weight_bytes"] = {"p4_codes+alpha": N * K // 2 + 4 * N, "fp16": 2 * N * K}
        out.append(r)
        print(json.dumps(r), flush=True)
json.dump({"peak_bf16_tflops": peak, "peak_bf16_tflops_sustained": peak_s, "cases": out},
          open(os.path.join(ROOT, "gpurun_out", "r02_gemm_bench.json"), "w"), indent=1)

"""BASELINE.json config C5, reproducible: synthetic N x N sweep (1k .. 16k) x numeric type (int / pot / flint /
float / OliVe int+abfloat / OliVe flint+abfloat) x scale granularity (per-tensor, per-row, group-8/16/32) x dtype,
kernel-only time of antq.fakequant (CUDA graph of `nb` launches over rotating buffer pairs larger than L2, CUDA events).

    python tools/sweep.py [--quick] [--out gpurun_out/r02_sweep.jsonl]

One JSON line per case: plan (which kernel), us per launch, algorithmic GB/s (sizeof(in) + sizeof(out) per element,
SURVEY.md 8(d)) and the fraction of the measured HBM peak (MEASURED_PEAKS.json, else the profiling guide's fallback).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200"))
import torch  # noqa: E402
import antq  # noqa: E402
from antq import _lib, codebooks  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def time_graph(step, reps):
    step(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def case(n, kind, bit, signed, olive, gran, dtype, flags=0, reps=20, alpha_scale=1.0, data="randn*0.02"):
    dev = torch.device("cuda:0")
    dt = {"f16": torch.float16, "f32": torch.float32, "bf16": torch.bfloat16}[dtype]
    es = 4 if dtype == "f32" else 2
    pair = n * n * es * 2
    nb = max(2, min(16, -(-(300 << 20) // pair)))                     # rotate > 126 MB L2
    if olive:
        cb = antq.prepare_codebook(codebooks.olive_grid(kind, bit, signed).to(dev), codebooks.olive_outliers(bit, signed).to(dev))
    else:
        cb = antq.prepare_codebook(codebooks.ant_grid(kind, bit, signed).to(dev))
    g = torch.Generator(device="cuda").manual_seed(n + bit)
    xs, als, outs = [], [], []
    for _ in range(nb):
        x = torch.randn(n, n, device=dev, generator=g)
        if data.startswith("tail"):                                   # 0.1 % of the entries x 20 / x 60
            x = torch.where(torch.rand(n, n, device=dev, generator=g) < 1e-3, x * float(data[4:]), x)
        elif data == "relu":                                          # post-ReLU, half the entries zero
            x = torch.relu(x)
        else:
            x = x * 0.02
            if not signed:
                x = x.abs()
        x = x.to(dt)
        if gran == "tensor":
            v, per_row = x, False
            al = x.float().abs().max().reshape(1) * 0.9
        elif gran.startswith("dyn"):                                  # dynamic group scales: abs-max inside the kernel
            v, per_row = x.view(-1, int(gran[3:])), True
            al = torch.zeros(1, device=dev)
        else:
            gsz = n if gran == "row" else int(gran[1:])
            v, per_row = x.view(-1, gsz), True
            al = v.float().abs().amax(1) * 0.9
        if olive:
            al = torch.full_like(al, float(3 * x.float().std()) * alpha_scale)
        xs.append(v); als.append(al.contiguous()); outs.append(torch.empty_like(v))
    plan = antq.fakequant_plan(xs[0], cb, per_row, ovp=olive, flags=flags)
    if gran.startswith("dyn"):
        plan = "dynamic (abs-max + fake-quant, one read)"

    def step():
        for i in range(nb):
            if gran.startswith("dyn"):
                antq.fakequant_dynamic(xs[i].view(-1), cb, int(gran[3:]), ratio=0.9, out=outs[i].view(-1))
            else:
                antq.fakequant(xs[i], als[i], cb, per_row, ovp=olive, out=outs[i], flags=flags)
    us = time_graph(step, reps) / nb
    pk, src = peak()
    gbs = n * n * es * 2 / us / 1e3
    return {"n": n, "type": ("olive-" if olive else "") + kind, "bit": bit, "signed": signed, "granularity": gran,
            "dtype": dtype, "data": data, "plan": plan, "us": round(us, 2), "GBps": round(gbs, 1), "frac": round(gbs / pk, 3),
            "peak": pk, "peak_source": src, "nb": nb}


def codes_case(n, kind, olive):
    dev = torch.device("cuda:0")
    nb = 6
    if olive:
        cb = antq.prepare_codebook(codebooks.olive_grid(kind, 4, True).to(dev), codebooks.olive_outliers(4, True).to(dev))
    else:
        cb = antq.prepare_codebook(codebooks.ant_grid(kind, 4, True).to(dev))
    g = torch.Generator(device="cuda").manual_seed(n)
    xs = [(torch.randn(n, n, device=dev, generator=g) * 0.02).to(torch.float16) for _ in range(nb)]
    als = [(x.float().abs().amax(1) * 0.9).contiguous() for x in xs]
    if olive:
        als = [torch.full_like(a, float(3 * x.float().std())) for a, x in zip(als, xs)]
    codes = [antq.encode_p4(x, a, cb, True, ovp=olive)[0] for x, a in zip(xs, als)]
    outs = [torch.empty_like(x) for x in xs]

    def enc():
        for i in range(nb):
            antq.encode_p4(xs[i], als[i], cb, True, ovp=olive, count_inexact=False)

    def dec():
        for i in range(nb):
            antq.decode_p4(codes[i], als[i], cb, xs[i].shape, torch.float16, True, ovp=olive, out=outs[i])
    pk, src = peak()
    r = {"n": n, "type": ("olive-" if olive else "") + kind, "bit": 4, "signed": True, "granularity": "row", "dtype": "f16",
         "plan": "packed 4-bit codes (P4)", "peak": pk, "peak_source": src}
    for name, fn in (("encode", enc), ("decode", dec)):
        us = time_graph(fn, 10) / nb
        gbs = n * n * 2.5 / us / 1e3
        r[name] = {"us": round(us, 2), "GBps": round(gbs, 1), "frac": round(gbs / pk, 3)}
    r["note"] = "2.5 algorithmic bytes per element (fp16 one way, 4-bit codes the other)"
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_sweep.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    sizes = [1024, 2048, 4096, 8192, 16384]
    types = [("int", 4, True, False), ("pot", 4, True, False), ("flint", 4, True, False), ("float2", 4, True, False),
             ("flint", 4, False, False), ("int", 4, False, False), ("int", 8, True, False), ("int", 8, False, False),
             ("int", 6, True, False), ("flint", 6, False, False), ("flint", 5, True, False),
             ("int", 4, True, True), ("flint", 4, True, True), ("flint", 4, False, True)]
    grans = ["tensor", "row", "g8", "g16", "g32", "g128"]
    rows = []
    with open(a.out, "w") as f:
        def emit(r):
            rows.append(r)
            f.write(json.dumps(r) + "\n"); f.flush()
            print(json.dumps(r), flush=True)
        if a.quick:
            for t in types:
                emit(case(4096, *t, "row", "f16"))
            return
        # every type x granularity at the headline size, fp16
        for t in types:
            for gr in grans:
                emit(case(4096, *t, gr, "f16"))
        # size sweep, per-row and per-tensor, the three headline types
        for n in sizes:
            if n == 4096:
                continue
            for t in (("flint", 4, True, False), ("int", 8, True, False), ("flint", 4, False, False), ("flint", 4, True, True)):
                for gr in ("row", "tensor", "g32"):
                    emit(case(n, *t, gr, "f16", reps=10 if n >= 8192 else 20))
        # dynamic group scales (single read) next to precomputed ones
        for t in (("flint", 4, True, False), ("int", 8, True, False), ("flint", 4, False, False)):
            for gr in ("dyn8", "dyn32", "dyn128"):
                emit(case(4096, *t, gr, "f16"))
        # dtypes
        for dt in ("f32", "bf16"):
            for t in (("flint", 4, True, False), ("int", 8, True, False), ("flint", 4, False, False)):
                for gr in ("row", "g32"):
                    emit(case(4096, *t, gr, dt))
        # the other synthetic inputs of SURVEY 8(d): heavy tails (clipped values, OliVe outliers), post-ReLU data
        for t, d in ((("flint", 4, True, False), "tail20"), (("flint", 4, True, False), "tail60"), (("int", 8, True, False), "tail20"),
                     (("flint", 4, True, True), "tail20"), (("flint", 4, True, True), "tail60"),
                     (("flint", 4, False, False), "relu"), (("int", 8, False, False), "relu"), (("flint", 4, False, True), "relu")):
            for gr in ("row", "tensor"):
                emit(case(4096, *t, gr, "f16", data=d))
        # int of every width
        for b in (3, 5, 7):
            for gr in ("row", "g32"):
                emit(case(4096, "int", b, True, False, gr, "f16"))
        # packed 4-bit codes: encode (2 B in + 0.5 B out per element) and decode (0.5 B in + 2 B out)
        emit(codes_case(4096, "flint", False))
        emit(codes_case(4096, "flint", True))
        # OliVe unsigned (post-ReLU activations, e.g. OPT's fc2 input): alpha = 3 std(x) of half-normal data leaves 4 % of the
        # elements beyond the first outlier threshold (the rows above); with alpha = 3 sigma of the underlying normal it is
        # 0.07 %, OliVe's design point -- and the other kernel (the two-phase chain) on the same data
        for fl, note in ((0, "alpha = 3 sigma: 0.07 % outliers"), (_lib.FLAG_NO_PU, "alpha = 3 sigma, NO_PU (two-phase chain)")):
            for gr in ("row", "tensor"):
                r = case(4096, "flint", 4, False, True, gr, "f16", flags=fl, alpha_scale=1.0 / 0.6028)
                r["note"] = note
                emit(r)
        # A/B: what the closed form replaced (chain / generic kernel on the same cases)
        for t in (("int", 8, True, False), ("flint", 4, False, False), ("int", 6, True, False), ("flint", 5, True, False)):
            for gr in ("row", "g32"):
                r = case(4096, *t, gr, "f16", flags=_lib.FLAG_NO_PU)
                r["note"] = "NO_PU (round-1 path)"
                emit(r)


if __name__ == "__main__":
    main()

"""Dequant-fused Linear on tcgen05 (antq_linear_p4) against F.linear on the fake-quantized operands
(A/antquant/quant_modules.py:642-646): a one-hot activation makes every output a single product, so the decode LUT, the
128-byte-swizzled operand layout, the UMMA descriptors and the tensor-memory epilogue are checked BIT-EXACTLY; random
activations are checked within the fp32-accumulate tolerance written below."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import antq_oracle as orc
from gpu_util import dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def antq():
    import antq as m
    return m


def _weights(antq, N, K, kind, signed, dtype, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = (torch.randn(N, K, generator=g) * 0.02).to(dtype).to(dev())
    if not signed:
        w = w.abs()
    grid = orc.ant_grid(kind, 4, signed)
    cb = antq.prepare_codebook(torch.from_numpy(grid).to(dev()))
    alpha = (w.float().abs().amax(1) * 0.9).contiguous()
    wq = antq.fakequant(w, alpha, cb, True)
    codes, bad = antq.encode_p4(w, alpha, cb, True)
    assert int(bad.item()) == 0
    return w, wq, codes, alpha, cb


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("kind,signed", [("flint", True), ("pot", True), ("flint", False)])
def test_linear_p4_one_hot_is_bit_exact(antq, kind, signed, dtype):
    N, K = 256, 256
    w, wq, codes, alpha, cb = _weights(antq, N, K, kind, signed, dtype, 1)
    x = torch.eye(K, dtype=dtype, device=dev())                        # y[m, n] = W_q[n, m]: one product per output
    y = antq.linear_p4(x, codes, alpha, cb, N)
    # every flint / pot level is exact in fp16 and bf16, so level * s rounds once on both sides
    assert torch.equal(y.view(torch.int16), wq.t().contiguous().view(torch.int16))
    bias = torch.randn(N, device=dev()).to(dtype)
    yb = antq.linear_p4(x, codes, alpha, cb, N, bias=bias)
    ref = (wq.t().float() + bias.float()).to(dtype)
    assert float((yb.float() - ref.float()).abs().max()) <= float(ref.float().abs().max()) * 2.0 ** -7


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 768, 512), (300, 256, 1024), (77, 512, 4096), (2048, 4096, 4096)])
def test_linear_p4_matches_f_linear(antq, M, N, K, dtype):
    w, wq, codes, alpha, cb = _weights(antq, N, K, "flint", True, dtype, 2)
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(M, K, generator=g).to(dtype).to(dev())
    bias = (torch.randn(N, generator=g) * 0.1).to(dtype).to(dev())
    y = antq.linear_p4(x, codes, alpha, cb, N, bias=bias)
    ref = F.linear(x.float(), wq.float(), bias.float())               # fp32 reference on the fake-quantized operands
    err = (y.float() - ref).norm() / ref.norm()
    # one rounding of the output to 16 bits (2^-11 fp16 / 2^-8 bf16 relative) dominates; fp32 accumulation is below it
    tol = 1.5e-3 if dtype == torch.float16 else 6e-3
    assert float(err) < tol, float(err)
    lib = F.linear(x, wq, bias)                                        # cuBLAS on the same operands, same dtype
    err_lib = (lib.float() - ref).norm() / ref.norm()
    assert float(err) < 2.0 * float(err_lib) + 1e-4, (float(err), float(err_lib))
    # 3-D activations keep their leading shape
    y3 = antq.linear_p4(x.view(1, M, K), codes, alpha, cb, N, bias=bias)
    assert y3.shape == (1, M, N) and torch.equal(y3.view(M, N), y)


def test_linear_p4_declines_what_it_cannot_do(antq):
    w, wq, codes, alpha, cb = _weights(antq, 256, 64, "flint", True, torch.float16, 4)
    x = torch.randn(8, 64, device=dev())
    with pytest.raises(RuntimeError):
        antq.linear_p4(x, codes, alpha, cb, 256)                       # fp32 activations: unsupported, loudly


def test_linear_quantizer_fused_path():
    """LinearQuantizer with antq.layers.FUSED_LINEAR: the weight is cached as packed codes and the layer output matches
    the unfused layer (fake-quantized fp16 weight + cuBLAS) to accumulation-order noise; training mode, fp32
    activations and OliVe outlier pairs keep the unfused path."""
    from host_util import run
    res, _ = run("ant", r'''
import antq.layers as L
dev = torch.device("cuda:0")
torch.manual_seed(0)
lin = nn.Linear(1024, 512).to(dev).half()
L.FUSED_MIN_ROWS = 64
q = LinearQuantizer(mode="flint", wbit=4, abit=4, args=mkargs("flint"))
q.set_param(lin)
q = q.to(dev).eval()
q.quant_weight.enable_quantization("w"); q.quant_input.enable_quantization("a")
x = torch.randn(4, 96, 1024, device=dev).half()
with torch.no_grad():
    y0 = q(x)                                    # calibrates, unfused
    y1 = q(x)
    L.FUSED_LINEAR = True
    y2 = q(x)
    pack = q._wc_val
    y3 = q(x)
    hit = q._wc_val is pack and pack is not None
    codes_bytes = pack[0].numel() if pack is not None else -1
    q.train(); yt = q(x); q.eval()                # training mode: unfused, cache dropped
    dropped = q._wc_val is None
    y4 = q(x.float().half())
    L.FUSED_LINEAR = False
L.FUSED_LINEAR = True; L.FUSED_FP8 = False
with torch.no_grad():
    y5 = q(x)                                    # 16-bit operand variant
L.FUSED_LINEAR = False; L.FUSED_FP8 = True
RESULT["rel16"] = float((y5.float() - y1.float()).norm() / y1.float().norm())
rel = float((y2.float() - y1.float()).norm() / y1.float().norm())
RESULT.update(rel=rel, hit=bool(hit), same=bool(torch.equal(y2, y3) and torch.equal(y3, y4)), codes_bytes=codes_bytes,
              train_unfused=bool(torch.equal(yt, y1)), dropped=bool(dropped), shape=list(y2.shape))
''', timeout=600)
    assert res["shape"] == [4, 96, 512] and res["hit"] and res["same"] and res["dropped"] and res["train_unfused"], res
    assert res["codes_bytes"] == 512 * 1024 // 2, res
    assert res["rel"] < 2e-3 and res["rel16"] < 2e-3, res


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("wkind,wsigned,xkind,xsigned", [("flint", True, "flint", False), ("int", True, "int", True), ("pot", True, "flint", True)])
@pytest.mark.parametrize("M,N,K", [(300, 256, 512), (2048, 1024, 4096)])
def test_linear_p4_fp8_is_exact_in_levels(antq, M, N, K, wkind, wsigned, xkind, xsigned, dtype):
    """W4A4 on the FP8 tensor cores: operands travel as e4m3 levels (small integers), so the accumulator is an exact
    integer and the output is reproducible bit for bit on the host: y = RN(fl32(acc * scale_n + bias_n))."""
    from antq import _lib
    w, wq, codes, alpha, cb = _weights(antq, N, K, wkind, wsigned, dtype, 5)
    xgrid = orc.ant_grid(xkind, 4, xsigned)
    xcb = antq.prepare_codebook(torch.from_numpy(xgrid).to(dev()))
    assert (cb.info.flags & _lib.CB_PU_E4M3) and (xcb.info.flags & _lib.CB_PU_E4M3)
    g = torch.Generator(device="cpu").manual_seed(6)
    x = torch.randn(M, K, generator=g).to(dtype).to(dev())
    if not xsigned:
        x = x.abs()
    xa = (x.float().abs().max() * 0.8).reshape(1)
    xq = antq.fakequant(x, xa, xcb, False)
    bias = (torch.randn(N, generator=g) * 0.1).to(dtype).to(dev())
    y = antq.linear_p4_fp8(xq, xa, xcb, codes, alpha, cb, N, bias=bias)
    # host replica: integer levels, exact integer GEMM in float64, the epilogue's fp32 operations in order
    f32 = np.float32
    cx = f32(xgrid[xgrid > 0].min())
    wgrid = orc.ant_grid(wkind, 4, wsigned)
    cw = f32(wgrid[wgrid > 0].min())
    sx = f32(f32(xa.item()) / f32(xgrid.max()))
    kx = np.rint(xq.float().cpu().numpy() / f32(sx * cx)).astype(np.float64)
    sw = (alpha.cpu().numpy().astype(f32) / f32(cb.info.gmax)).astype(f32)
    kw = np.rint(wq.float().cpu().numpy() / (sw[:, None] * cw)).astype(np.float64)
    assert np.abs(kx).max() <= 448 and np.abs(kw).max() <= 448
    acc = kx @ kw.T                                                       # exact: |acc| < 2^24
    assert np.abs(acc).max() < 2 ** 24
    unit = f32(f32(sx * cx) * cw)
    sc = (sw * unit).astype(f32)
    ref32 = (acc * sc[None, :].astype(np.float64) + bias.float().cpu().numpy().astype(np.float64)[None, :]).astype(f32)   # one rounding = fmaf
    ref = torch.from_numpy(ref32).to(dtype)
    assert torch.equal(y.cpu().view(torch.int16), ref.view(torch.int16)), int((y.cpu().view(torch.int16) != ref.view(torch.int16)).sum())
    # and it IS the layer's F.linear on the fake-quantized operands, up to their 16-bit roundings
    lib = F.linear(xq.float(), wq.float(), bias.float())
    err = (y.float() - lib).norm() / lib.norm()
    assert float(err) < (1.5e-3 if dtype == torch.float16 else 8e-3), float(err)

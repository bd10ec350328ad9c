"""Torch-facing wrappers over the C ABI.  Torch is plumbing only here: device
memory, the current stream, and a device guard; all arithmetic runs in libantq.so.

Reference functions replaced (A/ = ant_quantization/, O/ = olive_quantization/):
  lut_nearest   quant_cuda.quant                  A/quant/quant.cpp:26-28
  fakequant     Quantizer._forward                A/antquant/quant_modules.py:535-551
                OliVe _forward + OVP              O/antquant/quant_modules.py:295-330
  absmax        alpha init                        A/antquant/quant_modules.py:473-477
  mse_sweep     search_mse candidate loop         A/antquant/quant_modules.py:299-306
"""
import ctypes

import torch

from . import _lib
from ._lib import lib, check

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _dtype_code(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError("antq: unsupported dtype %s (float32, float16, bfloat16)" % t.dtype)


def _need_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("antq: %s must be a CUDA tensor -- there is no CPU path" % name)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class Codebook:
    """Device-resident prepared codebook + its host-side header."""

    def __init__(self, buf, info, k_normal, k_out):
        self.buf = buf
        self.info = info
        self.ptr = buf.data_ptr()
        self.info_ref = ctypes.byref(info)
        self.k_normal = k_normal
        self.k_out = k_out

    @property
    def device(self):
        return self.buf.device

    @property
    def n_entries(self):
        return self.info.n_entries

    def describe(self):
        return self.info.as_dict()


def prepare_codebook(grid, outliers=None):
    """grid / outliers: 1-D CUDA tensors (any float dtype; values are taken as fp32,
    like `quant_grid.type_as(x)` narrowed into the kernel's float smem table)."""
    _need_cuda(grid, "grid")
    with torch.cuda.device(grid.device):
        g = grid.detach().reshape(-1).to(torch.float32).contiguous()
        o = None
        if outliers is not None and outliers.numel() > 0:
            o = outliers.detach().reshape(-1).to(device=g.device, dtype=torch.float32).contiguous()
        k_out = 0 if o is None else o.numel()
        if g.numel() < 1 or g.numel() + k_out > _lib.MAX_GRID:
            raise ValueError("antq: grid must have 1..%d entries (got %d + %d)" % (_lib.MAX_GRID, g.numel(), k_out))
        buf = torch.empty(lib.antq_codebook_bytes(), dtype=torch.uint8, device=g.device)
        check(lib.antq_codebook_prepare(_ptr(g), g.numel(), _ptr(o), k_out, _ptr(buf), _stream()),
              "antq_codebook_prepare")
        info = _lib.CodebookInfo()
        check(lib.antq_codebook_info_get(_ptr(buf), ctypes.byref(info), _stream()), "antq_codebook_info_get")
        return Codebook(buf, info, g.numel(), k_out)


def lut_nearest(x, cb, want_codes=False):
    _need_cuda(x, "x")
    with torch.cuda.device(x.device):
        xc = x.contiguous()
        z = torch.empty_like(xc)
        codes = torch.empty(xc.shape, dtype=torch.int16, device=xc.device) if want_codes else None
        check(lib.antq_lut_nearest(_ptr(xc), _ptr(z), _ptr(codes), xc.numel(), _dtype_code(xc), _ptr(cb.buf),
                                   _stream()), "antq_lut_nearest")
    return (z, codes) if want_codes else z


def _rows_cols(x, per_row):
    if per_row:
        rows = x.shape[0] if x.dim() > 0 else 1
        return rows, (x.numel() // rows if rows else 0)
    return 1, x.numel()


def _alpha_arg(alpha, rows, per_row, device):
    a = alpha
    # only the pointer is read: a Parameter (requires_grad) is as good as its .detach()
    if not (a.dtype is torch.float32 and a.device == device and a.is_contiguous()):
        a = alpha.detach().to(device=device, dtype=torch.float32).contiguous()
    if a.numel() != (rows if per_row else 1):
        raise ValueError("antq: alpha has %d entries, expected %d" % (a.numel(), rows if per_row else 1))
    return a


def _check_out(out, x, name="out"):
    """A caller-supplied output buffer is handed to the kernel as a raw pointer: refuse anything that is not a
    same-sized, same-typed, contiguous tensor on x's device instead of writing out of bounds."""
    if not (isinstance(out, torch.Tensor) and out.dtype is x.dtype and out.device == x.device and
            out.numel() == x.numel() and out.is_contiguous()):
        raise ValueError("antq: `%s` must be a contiguous %s tensor with %d elements on %s" %
                         (name, x.dtype, x.numel(), x.device))


class _maybe_guard:
    """torch.cuda.device(...) only when the tensor is not on the current device (saves ~10 us per call)."""

    def __init__(self, device):
        self.g = None if device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.g is not None:
            self.g.__enter__()

    def __exit__(self, *exc):
        if self.g is not None:
            self.g.__exit__(*exc)


_raw_stream = torch._C._cuda_getCurrentRawStream       # (device index) -> cudaStream_t as int, no Stream object
_antq_fakequant = lib.antq_fakequant


def fakequant(x, alpha, cb, per_row, ovp=False, want_codes=False, flags=0, out=None):
    """Fused scale -> nearest -> (OVP) -> STE -> rescale.  x: contiguous CUDA tensor;
    alpha: fp32 CUDA tensor with x.shape[0] entries (per_row) or one entry.
    No allocation when `out` is given, no synchronisation: safe under CUDA-graph capture."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        _need_cuda(x, "x")
    if not x.is_contiguous():
        raise RuntimeError("antq: x must be contiguous")
    dev = x.device
    guard = None if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)
    if guard is not None:
        guard.__enter__()
    try:
        if per_row:
            rows = x.shape[0] if x.dim() > 0 else 1
            cols = x.numel() // rows if rows else 0
        else:
            rows, cols = 1, x.numel()
        a = _alpha_arg(alpha, rows, per_row, dev)
        if out is None:
            out = torch.empty_like(x)
        else:
            _check_out(out, x)
        codes = torch.empty(x.shape, dtype=torch.int16, device=dev) if want_codes else None
        fl = flags | (_lib.FLAG_OVP if ovp else 0)
        rc = _antq_fakequant(x.data_ptr(), out.data_ptr(), codes.data_ptr() if want_codes else None,
                             a.data_ptr(), 1 if per_row else 0, rows, cols, _dtype_code(x), cb.ptr,
                             cb.info_ref, fl, _raw_stream(dev.index))
        if rc:
            check(rc, "antq_fakequant")
    finally:
        if guard is not None:
            guard.__exit__(None, None, None)
    return (out, codes) if want_codes else out


def fakequant_grouped(x, alpha, cb, group_size, ovp=False, out=None):
    """Group-wise scales (one alpha per `group_size` consecutive elements of the flat tensor: group-8/16/32/128 ...):
    the tensor is viewed as [numel / group_size, group_size] and quantized per row.  Groups shorter than 512 elements
    take antq_short_kernel, longer ones the stream kernel.  alpha: fp32, numel / group_size entries."""
    _need_cuda(x, "x")
    if not x.is_contiguous():
        raise RuntimeError("antq: x must be contiguous")
    g = int(group_size)
    if g <= 0 or x.numel() % g:
        raise ValueError("antq: numel (%d) is not a multiple of the group size (%d)" % (x.numel(), g))
    xv = x.view(-1, g)
    if out is not None:
        _check_out(out, x)
    ov = None if out is None else out.view(-1, g)
    y = fakequant(xv, alpha, cb, True, ovp=ovp, out=ov)
    return y.view(x.shape) if out is None else out


def fakequant_dynamic(x, cb, group_size, ratio=1.0, return_alpha=False, out=None):
    """Group-wise DYNAMIC scales in one pass: alpha = max|x| over each `group_size` consecutive elements * ratio, then the
    fused fake-quant with that alpha (antq_fakequant_dynamic: one HBM read).  Falls back to absmax + fakequant_grouped
    (two reads) for the shapes / grids the single-pass kernel declines."""
    _need_cuda(x, "x")
    if not x.is_contiguous():
        raise RuntimeError("antq: x must be contiguous")
    g = int(group_size)
    if g <= 0 or x.numel() % g:
        raise ValueError("antq: numel (%d) is not a multiple of the group size (%d)" % (x.numel(), g))
    rows = x.numel() // g
    with _maybe_guard(x.device):
        if out is None:
            out = torch.empty_like(x)
        else:
            _check_out(out, x)
        alpha = torch.empty(rows, dtype=torch.float32, device=x.device) if return_alpha else None
        rc = lib.antq_fakequant_dynamic(_ptr(x), _ptr(out), _ptr(alpha), float(ratio), rows, g, _dtype_code(x), cb.ptr,
                                        cb.info_ref, 0, _stream())
        if rc == _lib.ENOTSUP:
            alpha = absmax(x.view(rows, g), True) * float(ratio)
            fakequant(x.view(rows, g), alpha, cb, True, out=out.view(rows, g))
        else:
            check(rc, "antq_fakequant_dynamic")
    return (out, alpha) if return_alpha else out


def fakequant_plan(x, cb, per_row, ovp=False, flags=0):
    rows, cols = _rows_cols(x, per_row)
    fl = flags | (_lib.FLAG_OVP if ovp else 0)
    return lib.antq_fakequant_plan(ctypes.byref(cb.info), rows, cols, _dtype_code(x), fl, _ptr(x), _ptr(x), None)


def absmax(x, per_row):
    _need_cuda(x, "x")
    with torch.cuda.device(x.device):
        xc = x.contiguous()
        rows, cols = _rows_cols(xc, per_row)
        out = torch.empty(rows, dtype=torch.float32, device=xc.device)
        check(lib.antq_absmax(_ptr(xc), _ptr(out), rows, cols, _dtype_code(xc), _stream()), "antq_absmax")
    return out


def mse_sweep(x, base_alpha, ratios, cb, per_row, ovp=False):
    """err[c, r] = sum_row (fakequant(x; alpha = base[r] * ratios[c]) - x)^2  (float64)."""
    _need_cuda(x, "x")
    with torch.cuda.device(x.device):
        xc = x.contiguous()
        rows, cols = _rows_cols(xc, per_row)
        a = _alpha_arg(base_alpha, rows, per_row, xc.device)
        r = ratios.detach().to(device=xc.device, dtype=torch.float32).reshape(-1).contiguous()
        err = torch.empty((r.numel(), rows), dtype=torch.float64, device=xc.device)
        check(lib.antq_mse_sweep(_ptr(xc), _ptr(a), int(bool(per_row)), _ptr(r), r.numel(), _ptr(err), rows, cols,
                                 _dtype_code(xc), _ptr(cb.buf), _lib.FLAG_OVP if ovp else 0, _stream()),
              "antq_mse_sweep")
    return err


def encode_p4(x, alpha, cb, per_row, ovp=False, count_inexact=True):
    """Packed 4-bit codes of the fake-quantized tensor (include/antq.h: antq_encode_p4).  Returns (codes, n_inexact):
    codes is a uint8 tensor of numel / 2 bytes (row-major, low nibble = even element); n_inexact a 1-element int32
    device tensor (None if not requested) counting the elements decode_p4 would not reproduce bit for bit."""
    _need_cuda(x, "x")
    if not x.is_contiguous():
        raise RuntimeError("antq: x must be contiguous")
    with _maybe_guard(x.device):
        rows, cols = _rows_cols(x, per_row)
        a = _alpha_arg(alpha, rows, per_row, x.device)
        codes = torch.empty(x.numel() // 2, dtype=torch.uint8, device=x.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=x.device) if count_inexact else None
        check(lib.antq_encode_p4(_ptr(x), _ptr(codes), _ptr(a), int(bool(per_row)), rows, cols, _dtype_code(x), cb.ptr,
                                 cb.info_ref, _lib.FLAG_OVP if ovp else 0, _ptr(cnt), _stream()), "antq_encode_p4")
    return codes, cnt


def decode_p4(codes, alpha, cb, shape, dtype, per_row, ovp=False, out=None):
    """values = level[code] * (alpha / max(grid)), rounded to `dtype` (antq_decode_p4)."""
    _need_cuda(codes, "codes")
    shape = tuple(shape)
    n = 1
    for d in shape:
        n *= d
    if codes.dtype is not torch.uint8 or codes.numel() * 2 != n or not codes.is_contiguous():
        raise ValueError("antq: codes must be a contiguous uint8 tensor of numel / 2 bytes")
    with _maybe_guard(codes.device):
        if out is None:
            out = torch.empty(shape, dtype=dtype, device=codes.device)
        elif not (out.dtype is dtype and out.device == codes.device and out.numel() == n and out.is_contiguous()):
            raise ValueError("antq: `out` must be a contiguous %s tensor with %d elements on %s" % (dtype, n, codes.device))
        rows, cols = _rows_cols(out.view(shape), per_row)
        a = _alpha_arg(alpha, rows, per_row, codes.device)
        check(lib.antq_decode_p4(_ptr(codes), _ptr(out), _ptr(a), int(bool(per_row)), rows, cols, _dtype_code(out), cb.ptr,
                                 cb.info_ref, _lib.FLAG_OVP if ovp else 0, _stream()), "antq_decode_p4")
    return out


def fakequant_backward(grad_out, x, out, alpha, gmax, per_row, need_grad_x=True, need_grad_alpha=True):
    """QAT backward of the fused forward in one pass (antq_fakequant_backward): returns (grad_x, grad_alpha[fp32])."""
    _need_cuda(grad_out, "grad_out")
    with _maybe_guard(x.device):
        g = grad_out if grad_out.is_contiguous() else grad_out.contiguous()
        if g.dtype is not x.dtype:
            g = g.to(x.dtype)
        xc = x if x.is_contiguous() else x.contiguous()
        oc = out if out.is_contiguous() else out.contiguous()
        rows, cols = _rows_cols(xc, per_row)
        a = _alpha_arg(alpha, rows, per_row, x.device)
        gx = torch.empty_like(xc) if need_grad_x else None
        ga = ws = None
        nws = 0
        if need_grad_alpha:
            ga = torch.empty(rows if per_row else 1, dtype=torch.float32, device=x.device)
            nws = lib.antq_backward_workspace_bytes(rows, cols, int(bool(per_row)))
            ws = torch.empty(max(nws, 8), dtype=torch.uint8, device=x.device)
        check(lib.antq_fakequant_backward(_ptr(g), _ptr(xc), _ptr(oc), _ptr(a), int(bool(per_row)), rows, cols,
                                          _dtype_code(xc), float(gmax), _ptr(gx), _ptr(ga), _ptr(ws), nws, _stream()),
              "antq_fakequant_backward")
    return gx, ga


def calibrate(x, base_alpha, ratios, cbs, per_row, ovp=False, want_index=False):
    """Fused calibration (antq_calibrate): every candidate alpha = base * ratio of every codebook in `cbs` scored in
    one read of x.  Returns (alpha[n_cb, rows], mse[n_cb][, best_index[n_cb, rows]]) -- device tensors, no host sync."""
    _need_cuda(x, "x")
    with _maybe_guard(x.device):
        xc = x if x.is_contiguous() else x.contiguous()
        rows, cols = _rows_cols(xc, per_row)
        base = _alpha_arg(base_alpha, rows, per_row, xc.device)
        r = ratios.detach().to(device=xc.device, dtype=torch.float32).reshape(-1).contiguous()
        n_cb, n_cand = len(cbs), r.numel()
        nrow = rows if per_row else 1
        alpha = torch.empty((n_cb, nrow), dtype=torch.float32, device=xc.device)
        mse = torch.empty(n_cb, dtype=torch.float32, device=xc.device)
        idx = torch.empty((n_cb, nrow), dtype=torch.int32, device=xc.device) if want_index else None
        nws = lib.antq_calibrate_workspace_bytes(rows, cols, int(bool(per_row)), n_cand, n_cb)
        ws = torch.empty(max(nws, 8), dtype=torch.uint8, device=xc.device)
        ptrs = (ctypes.c_void_p * n_cb)(*[c.ptr for c in cbs])
        infos = (ctypes.POINTER(_lib.CodebookInfo) * n_cb)(*[ctypes.pointer(c.info) for c in cbs])
        fl = (ctypes.c_int * n_cb)(*[(_lib.FLAG_OVP if ovp else 0)] * n_cb)
        check(lib.antq_calibrate(_ptr(xc), rows, cols, _dtype_code(xc), int(bool(per_row)), _ptr(base), _ptr(r), n_cand,
                                 ptrs, infos, fl, n_cb, _ptr(alpha), _ptr(mse), _ptr(idx), _ptr(ws), nws, _stream()),
              "antq_calibrate")
    return (alpha, mse, idx) if want_index else (alpha, mse)


def linear_p4(x, codes, alpha, cb, out_features, bias=None):
    """y = x . dequant(W)^T + bias on the tensor cores (antq_linear_p4): W is [out_features, in_features] held as packed
    4-bit codes + one alpha per output channel.  x: [..., in_features] fp16 / bf16 CUDA tensor."""
    _need_cuda(x, "x")
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M, N = x2.shape[0], int(out_features)
    with _maybe_guard(x.device):
        a = _alpha_arg(alpha, N, True, x.device)
        b = None
        if bias is not None:
            b = bias if (bias.dtype is x.dtype and bias.is_contiguous()) else bias.detach().to(x.dtype).contiguous()
        y = torch.empty((M, N), dtype=x.dtype, device=x.device)
        check(lib.antq_linear_p4(_ptr(x2), _ptr(codes), _ptr(a), _ptr(b), _ptr(y), M, N, K, _dtype_code(x2), cb.ptr,
                                 cb.info_ref, 0, _stream()), "antq_linear_p4")
    return y.view(*x.shape[:-1], N)


def linear_p4_fp8(x_q, x_alpha, x_cb, codes, alpha, cb, out_features, bias=None):
    """W4A4 on the FP8 tensor cores (antq_levels_e4m3 + antq_linear_p4_fp8): x_q is the fake-quantized activation
    (per-tensor alpha `x_alpha`, codebook `x_cb`); both operands travel as exact e4m3 levels, the scales are applied in the
    epilogue.  Needs ANTQ_CB_PU_E4M3 on both codebooks (every 4-bit int / flint / pot / float grid), K % 128 == 0,
    N % 256 == 0; raises otherwise."""
    _need_cuda(x_q, "x_q")
    K = x_q.shape[-1]
    x2 = x_q.reshape(-1, K)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M, N = x2.shape[0], int(out_features)
    with _maybe_guard(x_q.device):
        xa = _alpha_arg(x_alpha, 1, False, x_q.device)
        a = _alpha_arg(alpha, N, True, x_q.device)
        b = None
        if bias is not None:
            b = bias if (bias.dtype is x_q.dtype and bias.is_contiguous()) else bias.detach().to(x_q.dtype).contiguous()
        lev = torch.empty((M, K), dtype=torch.uint8, device=x_q.device)
        check(lib.antq_levels_e4m3(_ptr(x2), _ptr(lev), _ptr(xa), M * K, _dtype_code(x2), x_cb.ptr, x_cb.info_ref, _stream()),
              "antq_levels_e4m3")
        y = torch.empty((M, N), dtype=x_q.dtype, device=x_q.device)
        check(lib.antq_linear_p4_fp8(_ptr(lev), _ptr(xa), x_cb.ptr, x_cb.info_ref, _ptr(codes), _ptr(a), _ptr(b), _ptr(y),
                                     M, N, K, _dtype_code(x2), cb.ptr, cb.info_ref, 0, _stream()), "antq_linear_p4_fp8")
    return y.view(*x_q.shape[:-1], N)


class HostPipeline:
    """antq_host_*: fake-quant of HOST buffers (H2D, kernel, D2H pipelined in chunks)."""

    def __init__(self, device=0, chunk_bytes=8 << 20, n_stages=3):
        self._h = ctypes.c_void_p()
        self._keep = []
        check(lib.antq_host_create(ctypes.byref(self._h), int(device), int(chunk_bytes), int(n_stages)),
              "antq_host_create")

    def close(self):
        if self._h:
            lib.antq_host_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fakequant(self, x, out, alpha, grid, per_row, outliers=None, ovp=False, sync=True):
        """x/out: CPU tensors (ideally pinned), alpha/grid/outliers: CPU fp32 tensors.
        sync=False only enqueues the copies and kernels (antq_host_fakequant_async): `out` is complete after
        synchronize(), and consecutive tensors overlap on the PCIe link."""
        for t in (x, out, alpha, grid):
            if t.is_cuda:
                raise RuntimeError("HostPipeline takes host tensors")
        if not x.is_contiguous():
            raise RuntimeError("antq: x must be contiguous")
        _check_out(out, x)
        rows, cols = _rows_cols(x, per_row)
        a = alpha.detach().to(torch.float32).reshape(-1).contiguous()
        g = grid.detach().to(torch.float32).reshape(-1).contiguous()
        o = outliers.detach().to(torch.float32).reshape(-1).contiguous() if outliers is not None else None
        fn = lib.antq_host_fakequant if sync else lib.antq_host_fakequant_async
        check(fn(self._h, _ptr(x), _ptr(out), _ptr(a), int(bool(per_row)), rows, cols,
                 _dtype_code(x), _ptr(g), g.numel(), _ptr(o), 0 if o is None else o.numel(),
                 _lib.FLAG_OVP if ovp else 0), "antq_host_fakequant")
        if not sync:
            self._keep.append((x, out, a))           # the copies read / write these after the call returns
        return out

    def synchronize(self):
        check(lib.antq_host_synchronize(self._h), "antq_host_synchronize")
        del self._keep[:]

    @property
    def last_launches(self):
        return lib.antq_host_last_launches(self._h)

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
                             a.data_ptr(), 1 if per_row else 0, rows, cols, _DT[x.dtype], cb.ptr,
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

"""Harness that imports the UNMODIFIED reference Python from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/antq_oracle.c header).  It exists to
(1) validate the oracle restatement and (2) generate the golden fixtures in
tests/golden/ -- it only works in the build container, where /root/reference
is mounted; nothing that runs on the GPU box may import it.

The reference's native half (`quant_cuda`, A/quant/quant_kernel.cu) is CUDA
only, so a literal pure-torch restatement of its 27-line scan is injected as
module `quant_cuda`.  That stub is deliberately a
SECOND restatement, independent of oracle/antq_oracle.c, so the golden vectors
cross-check the C oracle instead of echoing it.
"""
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import torch

REF_ROOT = "/root/reference"
TREES = {
    "ant": os.path.join(REF_ROOT, "ant_quantization", "antquant"),
    "olive": os.path.join(REF_ROOT, "olive_quantization", "antquant"),
}


def reference_available():
    return os.path.isdir(TREES["ant"]) and os.path.isdir(TREES["olive"])


def _scan_stub(x, y):
    """A/quant/quant_kernel.cu:25-37 + launcher :48-61, as torch CPU ops."""
    xf = x.detach().to(torch.float32)          # `float x_v = x[idx]`
    yf = y.detach().to(torch.float32)          # `__shared__ float y_shared`
    best = torch.full_like(xf, 102400.0)
    z = torch.zeros_like(xf)
    for i in range(yf.numel()):
        sub = (xf - yf[i]).abs()
        take = sub <= best
        best = torch.where(take, sub, best)
        z = torch.where(take, yf[i], z)
    return z.to(x.dtype), torch.zeros_like(x)


def install_stub():
    mod = types.ModuleType("quant_cuda")
    mod.quant = _scan_stub
    sys.modules["quant_cuda"] = mod
    return mod


def ensure_gloo_group():
    """ANT's Quantizer calls dist.* unconditionally (A/antquant/quant_modules.py:517-531)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        import socket
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1)


def load_quant_modules(tree):
    """Import <tree>/antquant/quant_modules.py under a tree-specific name."""
    assert reference_available(), "/root/reference is not mounted"
    install_stub()
    name = "ref_%s_quant_modules" % tree
    if name in sys.modules:
        return sys.modules[name]
    path = TREES[tree]
    sys.path.insert(0, path)           # ANT star-imports quant_affine
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(path, "quant_modules.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(path)
        sys.modules.pop("quant_affine", None)
    if tree == "ant":
        ensure_gloo_group()
    return mod


def default_args(tree, **over):
    a = dict(w_up=150, a_up=150, w_low=75, a_low=75, percent=100, search=False)
    if tree == "olive":
        a.update(w_up=250, a_up=250, no_outlier=False)
    a.update(over)
    return SimpleNamespace(**a)


def make_quantizer(tree, mode, bit, is_signed, is_input, args=None, rows=None, name="t"):
    qm = load_quant_modules(tree)
    args = args or default_args(tree)
    q = qm.TensorQuantizer(mode=mode, bit=bit, is_signed=is_signed, is_enable=True,
                           is_input=is_input, args=args)
    if not is_input and rows is not None:
        q.alpha.data = torch.ones([rows, 1])       # mirrors set_param (:596,:634)
    q.enable_quantization(name)
    return q


def pin(q, grid, alpha, outliers=None):
    """Skip calibration: fix (grid, alpha)."""
    q.quant_grid.data = grid.clone()
    q.alpha.data = alpha.clone()
    if outliers is not None:
        q.outliers.data = outliers.clone()
    q.has_inited_quant_para.data = torch.tensor(1.0)
    return q


def grid_of(tree, kind, bit, is_signed, eb=None):
    """Run one of the reference's grid generators."""
    q = make_quantizer(tree, "int", bit, is_signed, is_input=True)
    if kind == "int":
        return q.int_value()
    if kind == "flint":
        return q.flint_value()
    if kind == "pot":
        return q.pot_value()
    if kind == "float":
        return q.float_value(eb if eb is not None else 3)
    if kind == "apot":
        return q.apot_value()
    if kind == "outlier":
        return q.outlier_value()
    raise ValueError(kind)

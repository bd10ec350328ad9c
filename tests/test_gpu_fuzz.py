"""Randomised shapes through the default plan against the generic kernel (which the oracle pins elsewhere): whatever kernel
`antq_fakequant_plan` picks for a (grid, dtype, rows, cols, scale granularity, pairs) combination must agree bit for bit
with ANTQ_FLAG_FORCE_FLAT.  Shapes are drawn to hit the seams: row lengths around the 512-element switch, one-vector rows,
rows that straddle tiles and chunks, partial last tiles, tensors smaller than one chunk, odd row counts."""
import numpy as np
import pytest
import torch

import antq_oracle as orc
from gpu_util import dev

pytestmark = pytest.mark.gpu

GRIDS = [("int", 8, True), ("int", 4, True), ("int", 3, False), ("flint", 4, True), ("flint", 4, False), ("flint", 6, False),
         ("pot", 4, False), ("float2", 4, True), ("apot", 4, False)]


def _cases(seed, n):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        vec_cols = int(rng.choice([1, 2, 3, 4, 7, 8, 16, 31, 32, 63, 64, 65, 100, 128, 257, 512, 1000]))
        rows = int(rng.choice([1, 2, 3, 5, 17, 64, 129, 333, 1024]))
        yield rng, vec_cols, rows


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
def test_default_plan_equals_generic_kernel(dtype):
    import antq
    from antq import _lib
    vec = 4 if dtype == torch.float32 else 8
    seen = set()
    for i, (rng, vec_cols, rows) in enumerate(_cases(11, 120)):
        kind, bit, signed = GRIDS[i % len(GRIDS)]
        cols = vec_cols * vec
        if rows * cols > (1 << 22):
            rows = max(1, (1 << 22) // cols)
        cb = antq.prepare_codebook(torch.from_numpy(orc.ant_grid(kind, bit, signed)).to(dev()))
        x = torch.from_numpy((rng.standard_normal((rows, cols)) * 0.02).astype(np.float32))
        x[rng.integers(0, rows), rng.integers(0, cols)] *= 50
        if rows > 3:
            x[2] = 0.0
            x[3, 0] = float("nan")
        if not signed:
            x = x.abs()
        x = x.to(dtype).to(dev())
        per_row = bool(rng.integers(0, 2)) and rows > 1
        if per_row:
            alpha = (x.float().nan_to_num().abs().amax(1) * float(rng.uniform(0.4, 1.1))).contiguous()
        else:
            alpha = (x.float().nan_to_num().abs().max() * 0.8).reshape(1)
        plan = antq.fakequant_plan(x, cb, per_row)
        seen.add(plan)
        y = antq.fakequant(x, alpha, cb, per_row)
        yf = antq.fakequant(x, alpha, cb, per_row, flags=_lib.FLAG_FORCE_FLAT)
        same = (y.view(torch.int16 if vec == 8 else torch.int32) == yf.view(torch.int16 if vec == 8 else torch.int32)) | (y.isnan() & yf.isnan())
        assert bool(same.all()), (kind, bit, signed, rows, cols, per_row, plan, int((~same).sum()))
    assert {1, 4, 5} <= seen or dtype == torch.float32, seen


@pytest.mark.parametrize("signed", [True, False])
def test_default_plan_equals_generic_kernel_olive(signed):
    import antq
    from antq import _lib
    g, o = orc.olive_grid("flint", 4, signed), orc.olive_outlier_grid(4, signed)
    cb = antq.prepare_codebook(torch.from_numpy(g).to(dev()), torch.from_numpy(o).to(dev()))
    for i, (rng, vec_cols, rows) in enumerate(_cases(23, 60)):
        cols = vec_cols * 8
        if rows * cols > (1 << 22):
            rows = max(1, (1 << 22) // cols)
        x = torch.from_numpy((rng.standard_normal((rows, cols)) * 0.02).astype(np.float32))
        x[torch.from_numpy(rng.random((rows, cols)) < 0.01)] *= 15
        if not signed:
            x = x.abs()
        x = x.to(torch.float16).to(dev())
        per_row = bool(rng.integers(0, 2)) and rows > 1
        alpha = torch.full((rows if per_row else 1,), 0.06, device=dev())
        y = antq.fakequant(x, alpha, cb, per_row, ovp=True)
        yf = antq.fakequant(x, alpha, cb, per_row, ovp=True, flags=_lib.FLAG_FORCE_FLAT)
        assert torch.equal(y.view(torch.int16), yf.view(torch.int16)), (signed, rows, cols, per_row, antq.fakequant_plan(x, cb, per_row, ovp=True))

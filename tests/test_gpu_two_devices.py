"""One process, two GPUs: every kernel family on a device that is NOT the current one.  The stream kernels opt in to
large dynamic shared memory with cudaFuncSetAttribute, which is per device: a process-wide "already configured" flag
would make the first launch on the second GPU fail.  Skipped on single-GPU boxes."""
import numpy as np
import pytest
import torch

import antq_oracle as orc
from gpu_util import assert_bit_equal, to_np

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_every_kernel_family_on_the_second_device():
    import antq
    from antq import _lib
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((64, 4096)) * 0.02).astype(np.float16)
    alpha = (np.abs(x.astype(np.float32)).max(1) * 0.9).astype(np.float32)
    for dev_id in (0, 1, 0):                                            # cuda:0 stays the current device throughout
        dev = torch.device("cuda", dev_id)
        assert torch.cuda.current_device() == 0
        xd, ad = torch.from_numpy(x).to(dev), torch.from_numpy(alpha).to(dev)
        for kind, bit, signed, plan in (("flint", 4, True, 1), ("int", 8, True, 4)):
            grid = orc.ant_grid(kind, bit, signed)
            cb = antq.prepare_codebook(torch.from_numpy(grid).to(dev))
            assert antq.fakequant_plan(xd, cb, True) == plan
            ref = orc.ant_forward(x, alpha, grid, per_row=True)
            assert_bit_equal(to_np(antq.fakequant(xd, ad, cb, True)), ref, "%s-%d on %s" % (kind, bit, dev))
            assert_bit_equal(to_np(antq.fakequant(xd, ad, cb, True, flags=_lib.FLAG_FORCE_FLAT)), ref, "flat on %s" % dev)
            xs = xd.view(-1, 32)
            a_s = (xs.float().abs().amax(1) * 0.9).contiguous()
            refs = orc.ant_forward(x.reshape(-1, 32), to_np(a_s), grid, per_row=True)
            assert_bit_equal(to_np(antq.fakequant(xs, a_s, cb, True)), refs, "short rows on %s" % dev)
        # OliVe pairs + the tcgen05 Linear (tensor map, TMEM) on the same device
        g, o = orc.olive_flint_grid(4, True), orc.olive_outlier_grid(4, True)
        cbo = antq.prepare_codebook(torch.from_numpy(g).to(dev), torch.from_numpy(o).to(dev))
        xo = x.astype(np.float32).reshape(-1)
        xo[::97] *= 40
        refo = orc.olive_forward(xo, np.float32(0.06), g, o, per_row=False)
        yo = antq.fakequant(torch.from_numpy(xo).to(dev), torch.tensor([0.06], device=dev), cbo, False, ovp=True)
        assert_bit_equal(to_np(yo), refo, "OVP on %s" % dev)
        cb4 = antq.prepare_codebook(torch.from_numpy(orc.ant_grid("flint", 4, True)).to(dev))
        w = xd[:, :256].contiguous().view(256, 64).repeat(1, 4).contiguous()          # [256, 256]
        aw = (w.float().abs().amax(1) * 0.9).contiguous()
        codes, bad = antq.encode_p4(w, aw, cb4, True)
        assert int(bad.item()) == 0
        wq = antq.fakequant(w, aw, cb4, True)
        yl = antq.linear_p4(torch.eye(256, dtype=torch.float16, device=dev), codes, aw, cb4, 256)
        assert torch.equal(yl, wq.t().contiguous()), "tcgen05 Linear on %s" % dev
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)

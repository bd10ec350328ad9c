"""The `Quantizer` / `TensorQuantizer` of antquant, backed by the fused sm_100a kernels.

Mirror of the reference worker classes (same constructor, attributes, buffers and
method names, so `quant_model.py` / `quant_utils.py` style code and checkpoints keep
working) with the arithmetic moved into libantq.so:

  reference                                             here
  QuantBase._quantization + quant_cuda.quant            antq.lut_nearest            (A/antquant/quant_modules.py:11-24)
  Quantizer._forward (7 elementwise passes + scan)      ONE antq_fakequant launch   (A/...:535-551, O/...:295-330)
  search_mse (75-88 x ~20 passes)                       ONE antq_mse_sweep launch   (A/...:287-326, O/...:190-233)
  abs-max init                                          antq_absmax                 (A/...:473-477)

Two flavours share the class: "ant" (A/antquant/quant_modules.py) and "olive"
(O/antquant/quant_modules.py: threshold-32 grids, abfloat outliers, 3-sigma alpha
init, outlier-victim pairs, everything under no_grad).

There is no CPU arithmetic path: a quantizer that is enabled raises on CPU tensors.
"""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import codebooks, ops


def _dist_on():
    return dist.is_available() and dist.is_initialized()


def _rank0():
    return (not _dist_on()) or dist.get_rank() == 0


class _FakeQuantSTE(torch.autograd.Function):
    """out = ((q - d).detach() + d) * s with d = x / s, s = alpha / max(grid)  (A/...:535-551).
    d out/d x = 1 (also for clipped elements); d out/d alpha = sum(g * (q - d)) / max(grid)."""

    @staticmethod
    def forward(ctx, x, alpha, quantizer):
        out = quantizer._launch(x, alpha)
        ctx.quantizer = quantizer
        ctx.alpha_shape = alpha.shape
        ctx.save_for_backward(x, out, alpha)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out, alpha = ctx.saved_tensors
        q = ctx.quantizer
        need_x, need_a = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_x or need_a):
            return None, None, None
        # one fused pass (antq_fakequant_backward): grad_x = fl(fl(g * s) / s), exactly what the reference's autograd
        # produces (mul then div); grad_alpha = sum_row g * (q - d) / max(grid), reduced in a fixed order
        gx, ga = ops.fakequant_backward(g, x, out, alpha, q._grid_max_host(), q.is_perchannel, need_x, need_a)
        grad_x = gx.view(g.shape) if need_x else None
        grad_alpha = ga.reshape(ctx.alpha_shape).to(alpha.dtype) if need_a else None
        return grad_x, grad_alpha, None


class QuantBase:
    """quant_cuda.quant behind the reference's static helper (A/antquant/quant_modules.py:11-24)."""

    @staticmethod
    def _quantization(x, quant_grid):
        cb = ops.prepare_codebook(quant_grid.to(x.device))
        flat = x.reshape(-1)
        if flat.dtype == torch.float64:                      # the kernel narrows doubles to float (quant_kernel.cu:28)
            return ops.lut_nearest(flat.float().contiguous(), cb).double().view(x.shape)
        return ops.lut_nearest(flat.contiguous(), cb).view(x.shape)

    @staticmethod
    def forward(real_val, quant_grid):
        with torch.no_grad():
            return QuantBase._quantization(real_val, quant_grid)


# Activation quantizers with identical parameters fed the SAME tensor object (q / k / v projections of an attention block all
# quantize the block's input, each with its own -- identically calibrated -- quantizer: the reference launches three times)
# share one launch in no-grad mode.  The result is the same tensor bit for bit; ANTQ_SHARE_INPUT_QUANT=0 turns it off.
SHARE_INPUT_QUANT = os.environ.get("ANTQ_SHARE_INPUT_QUANT", "1") != "0"
_capturing = getattr(torch._C, "_cuda_isCurrentStreamCapturing", None) or torch.cuda.is_current_stream_capturing


class Quantizer(nn.Module):
    flavor = "ant"
    _share_ids = {}                  # (grid, outliers, alpha, pairs) -> small int
    _share_memo = None               # (x, x._version, share id, out, out._version, ids of the quantizers served, capturing?)

    def __init__(self, mode="base", bit=8, is_signed=True, is_enable=False, is_input=False, args=None, operator=None):
        super().__init__()
        self.mode = mode
        self.is_input = is_input
        self.is_signed = is_signed
        self.is_enable = is_enable
        self.is_enable_activation = is_enable
        self.is_enable_weight = is_enable
        self.args = args
        self.operator = operator

        self.alpha = nn.Parameter(torch.tensor(1.0, requires_grad=True))
        self.register_buffer('bit', torch.tensor(bit))
        self.register_buffer('has_inited_quant_para', torch.tensor(0.0))
        self.register_buffer('quant_grid', torch.ones(2 ** bit))
        if self.flavor == "olive":
            self.register_buffer('outliers', torch.ones(2 ** bit))

        self.w_up, self.a_up = self.args.w_up, self.args.a_up
        self.w_low, self.a_low = self.args.w_low, self.args.a_low
        self.percent = self.args.percent / 100
        self.is_perchannel = not is_input            # inputs are never per-channel
        self.search = args.search
        self.mse = torch.tensor(0.0)
        self.name = None

        self._cb = None
        self._cb_key = None
        self._cb_src = None
        self._ovp = self.flavor == "olive" and not bool(getattr(self.args, "no_outlier", False))
        self._inited_key = None
        self._inited_val = False
        self._share = None                           # (share id, alpha version, grid version) once the parameters are known
        self._share_tried = None

    # ------------------------------------------------------------------ toggles
    def disable_input_quantization(self):
        self.is_enable_activation = False

    def enable_quantization(self, name):
        self.name = name
        self.is_enable = True

    def disable_quantization(self, name):
        self.name = name
        self.is_enable = False

    def update_signed(self, tensor):
        if not self.is_signed and tensor.min() < 0:        # (already signed: nothing to learn, no host round trip)
            self.is_signed = True

    # ---------------------------------------------------------------- codebooks
    def _bits(self):
        return int(self.bit.item())

    def _no_outlier(self):
        return self.flavor != "olive" or bool(getattr(self.args, "no_outlier", False))

    def convert_tensor(self, values):
        return codebooks._ant_finish(list(values), self._bits(), self.quant_grid.device)

    def int_value(self, q_type="int"):
        if self.flavor == "olive":
            return codebooks.olive_grid("int", self._bits(), self.is_signed, self.quant_grid.device)
        if q_type == "int":
            return codebooks.ant_grid("int", self._bits(), self.is_signed, self.quant_grid.device)
        B = codebooks._vbits(self._bits(), self.is_signed)
        return self.convert_tensor(codebooks._mirror(codebooks.int_magnitudes(B), self.is_signed))

    def flint_value(self, exp_base=0):
        if self.flavor == "olive":
            return codebooks.olive_grid("flint", self._bits(), self.is_signed, self.quant_grid.device)
        g = codebooks.ant_grid("flint", self._bits(), self.is_signed, self.quant_grid.device)
        return g          # exp_base only shifts the table before it is renormalised to max = 10

    def pot_value(self):
        return codebooks.ant_grid("pot", self._bits(), self.is_signed, self.quant_grid.device)

    def float_value(self, eb=3):
        return codebooks.ant_grid("float%d" % eb, self._bits(), self.is_signed, self.quant_grid.device)

    def apot_value(self):
        return codebooks.ant_grid("apot", self._bits(), self.is_signed, self.quant_grid.device)

    def outlier_value(self, exp_bit=2, exp_base=5):
        return codebooks.olive_outliers(self._bits(), self.is_signed, exp_bit, exp_base, self.quant_grid.device)

    def _grid_for(self, kind):
        if self.flavor == "olive":
            if kind in ("int", "flint"):
                return codebooks.olive_grid(kind, self._bits(), self.is_signed, self.quant_grid.device)
            raise RuntimeError("Unsupported mode: " + kind)
        return codebooks.ant_grid(kind, self._bits(), self.is_signed, self.quant_grid.device)

    def _grid_max(self):
        return torch.max(self.quant_grid)

    def _grid_max_host(self):
        """max(quant_grid) as a Python float without a device round trip: the prepared codebook's header has it."""
        return float(self._codebook(self.quant_grid.device).info.gmax) if self.quant_grid.is_cuda else float(self.quant_grid.max())

    def _codebook(self, device):
        """Prepared device codebook, rebuilt only when the grid buffers change
        (load_state_dict / load_ant_state_dict / type search assign new tensors).  When only the VERSION of a
        buffer moved (DDP's broadcast_buffers copies rank 0's buffers in place before every forward) the contents
        are compared on the device first, so an unchanged grid costs one tiny compare instead of a rebuild."""
        g = self.quant_grid
        o = None if self._no_outlier() else self.outliers
        key = (g.data_ptr(), g._version, g.numel(), device,
               None if o is None else (o.data_ptr(), o._version, o.numel()))
        if key != self._cb_key:
            old = self._cb_key
            same = False
            if old is not None and self._cb_src is not None and old[0] == key[0] and old[2:4] == key[2:4] and \
                    (o is None) == (old[4] is None) and (o is None or (old[4][0], old[4][2]) == (key[4][0], key[4][2])):
                same = bool(torch.equal(g, self._cb_src[0]) and (o is None or torch.equal(o, self._cb_src[1])))
            if not same:
                self._cb = ops.prepare_codebook(g.to(device), None if o is None else o.to(device))
                self._cb_src = (g.detach().clone(), None if o is None else o.detach().clone())
            self._cb_key = key
        return self._cb

    # ------------------------------------------------------------------ forward
    def _make_share_key(self):
        """Host-side identity of this quantizer's parameters (one device read: called where calibration has just
        synchronised anyway).  Only per-tensor activation quantizers share launches.  The key is dropped when alpha or the
        grid change version (in-place updates); code that swaps their storage (`.data = ...`) calls this again or
        `invalidate_share()`."""
        self._share = None
        if not (self.is_input and not self.is_perchannel and self.alpha.numel() == 1 and self.alpha.is_cuda):
            return
        key = (tuple(self.quant_grid.flatten().tolist()), tuple(self.outliers.flatten().tolist()) if self._ovp else (),
               float(self.alpha), bool(self._ovp), str(self.alpha.device))
        sid = Quantizer._share_ids.setdefault(key, len(Quantizer._share_ids))
        self._share = (sid, self.alpha._version, self.quant_grid._version)

    def invalidate_share(self):
        self._share = None

    def _launch(self, x, alpha):
        if not x.is_cuda:
            raise RuntimeError("antquant (B200): quantization needs CUDA tensors; there is no CPU fallback")
        sh = self._share
        if sh is None and self.is_input and SHARE_INPUT_QUANT and not torch.is_grad_enabled():
            # parameters that did not come from a calibration in this process (a loaded checkpoint, an in-place update):
            # read them once per version, outside graph capture
            tried = (self.alpha._version, self.quant_grid._version)
            if self._share_tried != tried and self._is_inited() and not torch.cuda.is_current_stream_capturing():
                self._share_tried = tried
                self._make_share_key()
                sh = self._share
        if sh is not None and SHARE_INPUT_QUANT and alpha is self.alpha and not torch.is_grad_enabled():
            if sh[1] != alpha._version or sh[2] != self.quant_grid._version:
                self._share = None                    # parameters were touched since: no sharing until re-calibrated
            else:
                # A hit needs the same tensor OBJECT at the same version, the same parameters, an untouched result -- and a
                # quantizer that has not used this entry yet: the same quantizer coming back means a new forward pass (or
                # a CUDA-graph capture after its warm-up), which must launch again.
                # ... and the same capture state: a result recorded while a CUDA graph was being captured has not been
                # computed yet, and an eager result is not part of a graph being captured now.
                m = Quantizer._share_memo
                cap = _capturing()
                if (m is not None and m[0] is x and m[1] == x._version and m[2] == sh[0] and m[4] == m[3]._version
                        and id(self) not in m[5] and m[6] == cap):
                    m[5].add(id(self))
                    return m[3]
                if x.is_contiguous():
                    out = ops.fakequant(x, alpha, self._codebook(x.device), self.is_perchannel, self._ovp)
                else:
                    xc = x.contiguous()
                    out = ops.fakequant(xc, alpha, self._codebook(xc.device), self.is_perchannel, self._ovp).view(x.shape)
                Quantizer._share_memo = (x, x._version, sh[0], out, out._version, {id(self)}, cap)
                return out
        if x.is_contiguous():
            return ops.fakequant(x, alpha, self._codebook(x.device), self.is_perchannel, self._ovp)
        xc = x.contiguous()
        return ops.fakequant(xc, alpha, self._codebook(xc.device), self.is_perchannel, self._ovp).view(x.shape)

    def _forward(self, data, display=False):
        if self.flavor == "olive" or not torch.is_grad_enabled() or not (data.requires_grad or self.alpha.requires_grad):
            return self._launch(data, self.alpha)         # the kernel only reads alpha's storage: no graph is recorded
        return _FakeQuantSTE.apply(data, self.alpha, self)

    # -------------------------------------------------------------- calibration
    def mse_loss(self, quant_tensor, source_tensor, p=2.0, is_perchannel=True):
        d = (quant_tensor - source_tensor).abs().pow(p)
        if is_perchannel:
            return d.view(quant_tensor.shape[0], -1).mean(-1).unsqueeze(1)
        return d.mean()

    def _base_alpha(self, tensor, per_row):
        if self.flavor == "olive" and not self._no_outlier():        # 3-sigma clipping init (O/...:192-198,213-218)
            if per_row:
                v = tensor.reshape(tensor.shape[0], -1).float()
                mean, std = v.mean(dim=-1), v.std(dim=-1)
            else:
                v = tensor.float()
                mean, std = v.mean(), v.std()
            return torch.maximum((mean + 3 * std).abs(), (mean - 3 * std).abs()).reshape(-1)
        return ops.absmax(tensor, per_row)

    def _candidates(self, per_row):
        lb, ub = (int(self.w_low), int(self.w_up)) if per_row else (int(self.a_low), int(self.a_up))
        if self.flavor == "ant":
            if self.bit > 6:
                lb = 95
            return [i * 0.01 for i in range(lb, ub)]
        return [i * 0.01 for i in range(lb, ub, 2)]

    def _calib_inputs(self, tensor):
        per_row = self.is_perchannel and (not self.is_input)
        x = tensor.detach()
        x = x if x.is_contiguous() else x.contiguous()
        base = self._base_alpha(x, per_row).float()
        ratios = torch.tensor(self._candidates(per_row), dtype=torch.float32, device=x.device)
        return per_row, x, base, ratios

    _KIND_CODEBOOKS = {}

    def _cb_of_kind(self, kind, device):
        """Prepared codebook of a grid family: a pure function of (flavour, kind, bits, sign, outliers, device), so it
        is built once per process -- the type search then costs no codebook build and no host round trip."""
        key = (self.flavor, kind, self._bits(), bool(self.is_signed), self._no_outlier(), str(device))
        cb = Quantizer._KIND_CODEBOOKS.get(key)
        if cb is None:
            o = None if self._no_outlier() else self.outlier_value().to(device)
            cb = ops.prepare_codebook(self._grid_for(kind).to(device), o)
            Quantizer._KIND_CODEBOOKS[key] = cb
        return cb

    def _score_grids(self, tensor, cbs):
        """(alpha[n, rows], score[n]) for a list of prepared codebooks: ONE fused launch (antq_calibrate) that reads
        the tensor once and scores every (grid, alpha candidate) pair; device tensors, no host synchronisation."""
        per_row, x, base, ratios = self._calib_inputs(tensor)
        ovp = not self._no_outlier()
        cols = x.numel() // base.numel()
        if ovp and (x.numel() % 2 or (per_row and cols % 2)):
            # pairs straddle rows / wrap around: rare shapes, candidate loop through the fused forward
            alphas, scores = [], []
            for cb in cbs:
                err = self._sweep_loop(x, base, ratios, per_row, cb) / cols
                best, _ = err.min(dim=0)
                first = (err == best.unsqueeze(0)).to(torch.int8).argmax(dim=0)
                alphas.append(base * ratios[first]); scores.append(best.sum().float())
            return torch.stack(alphas), torch.stack(scores), per_row, base
        alpha, score = ops.calibrate(x, base, ratios, cbs, per_row, ovp=ovp)
        return alpha, score, per_row, base

    def _shape_alpha(self, a, per_row):
        return a.unsqueeze(1) if per_row else a.reshape(())

    def search_mse(self, tensor):
        """One fused sweep instead of the reference's Python loop: every candidate
        alpha = base * (i * 0.01) is scored in a single pass over the tensor (A/...:287-326, O/...:190-233)."""
        alpha, score, per_row, base = self._score_grids(tensor, [self._codebook(tensor.device)])
        a = self._shape_alpha(alpha[0], per_row)
        self.alpha.data = a.to(self.alpha.dtype)
        return score[0], a, (alpha[0] / base).mean()

    def _sweep_loop(self, x, base, ratios, per_row, cb=None):
        """Candidate loop through the fused forward (shapes the fused calibration declines)."""
        cb = cb or self._codebook(x.device)
        errs = []
        for r in ratios:
            a = (base * r)
            q = ops.fakequant(x, a, cb, per_row, not self._no_outlier())
            e = (q.double() - x.double()) ** 2
            errs.append(e.reshape(base.numel(), -1).sum(1))
        return torch.stack(errs)

    def _type_candidates(self):
        """(token, grid kind the PROBE scores, grid kind the winner gets) in the reference's order
        (A/...:328-415; OliVe: int and flint only, O/...:236-256)."""
        mode = self.mode
        order = ["int", "flint"] if self.flavor == "olive" else \
            ["int", "flint", "pot", "float", "float1", "float2", "float3", "float4", "apot"]
        out = []
        for tok in order:
            if ("-" + tok) not in mode:
                continue
            # reference quirk kept: the -float2/3/4 probes all score float_value(1) (A/...:379,388,397)
            out.append((tok, "float1" if tok in ("float2", "float3", "float4") else tok, tok))
        return out

    def search_adaptive_numeric_type(self, data):
        """One type per tensor: the candidate whose best summed MSE is smallest.  All candidate types (and the winner's
        final grid) are scored by ONE launch; the only host read is the handful of per-type scores."""
        cands = self._type_candidates()
        kinds = []
        for _, probe, final in cands:
            for k in (probe, final):
                if k not in kinds:
                    kinds.append(k)
        alpha, score, per_row, _ = self._score_grids(data, [self._cb_of_kind(k, data.device) for k in kinds])
        host = score.cpu().numpy()                                      # the one synchronisation of the type search
        probe_scores = np.array([host[kinds.index(p)] for _, p, _ in cands])
        win = cands[int(np.argsort(probe_scores)[0])]
        self.mode = win[0]
        self._calibrated = (self._grid_for(win[2]), self._shape_alpha(alpha[kinds.index(win[2])], per_row))

    def outlier_set(self, data):
        """OLAccel-style baseline (mode == 'outlier', A/...:417-436): the int-4 window ends at the `percent`
        percentile of |x| (np.percentile on the host, like the reference), the 16-bit window at max |x|."""
        def reduce_ave(t):
            rt = t.clone()
            if _dist_on():
                dist.all_reduce(rt, op=dist.ReduceOp.SUM)
                rt /= dist.get_world_size()
            return rt
        host = data.abs().cpu()
        if host.dtype in (torch.float16, torch.bfloat16):
            host = host.float()
        self.percent_value_int4 = torch.tensor(np.percentile(host.numpy(), self.percent * 100), device=data.device)
        self.percent_value_int16 = data.abs().max()
        self.percent_value_int4.data = reduce_ave(self.percent_value_int4.data)
        self.percent_value_int16.data = reduce_ave(self.percent_value_int16.data)
        if _rank0():
            print(self.name, self.percent_value_int4.item(), self.percent_value_int16.item())
        self.is_perchannel = False
        self.quant_grid.data = self.int_value()
        self.has_inited_quant_para.data = torch.ones_like(self.has_inited_quant_para)

    def outlier_quant(self, data):
        """A/...:438-465, op for op: the int-4 part is `nearest(data / scale) * scale` -- no STE sum, unlike
        `_forward` -- so it goes through antq_lut_nearest (the scan kernel's drop-in), not the fused kernel."""
        p4, p16 = self.percent_value_int4, self.percent_value_int16
        mask_int16 = data.abs() > p4
        if p4 > 0:
            scale = p4 / torch.max(self.quant_grid)
            data_int4 = (data / scale).detach()
            quant_data = ops.lut_nearest(data_int4.contiguous(), self._codebook(data.device)).view(data.shape)
            tensor = quant_data * scale
        else:
            tensor = data.clone().detach()
        level = 2 ** 16 - 1 if self.is_signed else 2 ** 15 - 1
        if self.percent < 100:
            scale = (p16 - p4) / level
            data_int16 = data[mask_int16].abs()
            sign_int16 = data[mask_int16].sign()
            data_int16 = data_int16 - p4
            quant_data = (data_int16 / scale).round() * scale
            quant_data = quant_data + p4
            quant_data = quant_data * sign_int16
            tensor[mask_int16] = (quant_data - tensor[mask_int16]).detach() + tensor[mask_int16]
        return tensor

    def _is_inited(self):
        """`has_inited_quant_para == 0` without a host sync on every call: the buffer is re-read
        only when it was reassigned (load_state_dict, set_8_bit_layer_*)."""
        b = self.has_inited_quant_para
        key = (b.data_ptr(), b._version)
        if key != self._inited_key:
            self._inited_val = bool(b.item() != 0)
            self._inited_key = key
        return self._inited_val

    def _init_quant_para(self, data, data_b=None):
        with torch.no_grad():
            if self._is_inited():
                return
            self.update_signed(data)
            if self.flavor == "olive":
                self.outliers.data = self.outlier_value().to(self.outliers.device)
            per_row = self.is_perchannel
            a0 = ops.absmax(data.detach().contiguous(), per_row)
            self.alpha.data = a0.unsqueeze(1) if per_row else a0.reshape(())

            if self.flavor == "ant" and self.mode == 'outlier':
                return self.outlier_set(data)

            self._calibrated = None
            if self.bit > 6:
                self.mode = 'int'
            elif "ant-" in self.mode:
                self.search_adaptive_numeric_type(data)

            valid = ("int", "flint") if self.flavor == "olive" else \
                ("int", "flint", "pot", "apot", "float", "float1", "float2", "float3", "float4")
            if self.mode not in valid:
                raise RuntimeError("Unsupported mode: " + self.mode)
            if self._calibrated is not None:
                # the type search already scored the winner's own grid: the final search_mse of the reference
                # (A/...:513) would repeat exactly that computation
                grid, alpha = self._calibrated
                self.quant_grid.data = grid
                self._calibrated = None
            else:
                self.quant_grid.data = self._grid_for(self.mode)
                _, alpha, _ = self.search_mse(data)
            self.alpha.data = alpha.to(self.alpha.dtype)

            quant_data = self._forward(data)
            self.mse = self.mse_loss(quant_data.float(), data.float(), 2, is_perchannel=self.is_perchannel).mean()
            if self.flavor == "ant" and _dist_on():
                dist.broadcast(self.mse, 0)
            if _rank0() or self.flavor == "olive":
                print(self.mode, end="\t")
                print("%d-bit \t %s," % (self.bit.item(), self.name))
            self._sync_after_calibration()
            self.has_inited_quant_para.data = torch.ones_like(self.has_inited_quant_para)
            self._make_share_key()

    def _sync_after_calibration(self):
        """The only exchange step of the path, once per quantizer: the calibrated scale is averaged
        over the data-parallel ranks and rank 0's type choice (its grid) wins (A/...:517-531).
        No-op without a process group and for OliVe (whose reference never calls torch.distributed)."""
        if self.flavor != "ant" or not _dist_on():
            return
        rt = self.alpha.data.clone()
        dist.all_reduce(rt, op=dist.ReduceOp.SUM)
        self.alpha.data = rt / dist.get_world_size()
        dist.broadcast(self.quant_grid, 0)

    def tensor_forward(self, tensor, input_tensor=None):
        if self.mode == "base" or not self.is_enable:
            return tensor
        if self.is_input:
            if not self.is_enable_activation:
                return tensor
        elif not self.is_enable_weight:
            return tensor
        if not self._is_inited():
            with torch.no_grad():
                self._init_quant_para(tensor, input_tensor)
        if self.flavor == "ant" and self.mode == 'outlier':
            return self.outlier_quant(tensor)
        return self._forward(tensor)


class TensorQuantizer(Quantizer):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def forward(self, tensor, input_tensor=None):
        return self.tensor_forward(tensor, input_tensor)


class OliveQuantizer(Quantizer):
    flavor = "olive"

    @torch.no_grad()
    def tensor_forward(self, tensor, input_tensor=None):
        return super().tensor_forward(tensor, input_tensor)


class OliveTensorQuantizer(OliveQuantizer):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def forward(self, tensor, input_tensor=None):
        return self.tensor_forward(tensor, input_tensor)

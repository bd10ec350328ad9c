"""antquant.quant_utils: the global quantizer configuration and model-wide toggles
(same names as ant_quantization/antquant/quant_utils.py; `from quant_utils import *`
also leaks os, torch, logging, uuid, models, dist, which the reference drivers use)."""
import _bootstrap  # noqa: F401
import os
import torch
import logging
from quant_modules import Quantizer as Q
import uuid
import torch.distributed as dist

try:
    import torchvision.models as models
except Exception:        # torchvision is only needed by get_model()
    models = None

quant_args = {}
logger = logging.getLogger(__name__)


def set_quantizer(args):
    """Every Quantizer created afterwards receives mode / wbit / abit and the whole namespace."""
    quant_args.update(mode=args.mode, wbit=args.wbit, abit=args.abit, args=args)


def set_util_logging(filename):
    logging.basicConfig(format='%(asctime)s - %(levelname)s - %(name)s -   %(message)s',
                        datefmt='%m/%d/%Y %H:%M:%S', level=logging.INFO,
                        handlers=[logging.FileHandler(filename), logging.StreamHandler()])


def tag_info(args):
    return "_" + args.tag if args.tag != "" else ""


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_ckpt_path(args):
    """output/<model>_<dataset>/<mode>_W<w>A<a>_<id>/gpu_<rank>; rank 0 draws the id and broadcasts it."""
    rank = _rank()
    run = int(uuid.uuid4().hex[0:4], 16)
    if dist.is_available() and dist.is_initialized():
        dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        t = torch.tensor(run, device=dev)
        dist.broadcast(t, 0)
        run = int(t.item())
    base = os.path.join('output', args.model + "_" + args.dataset,
                        "%s_W%sA%s_%d" % (args.mode, args.wbit, args.abit, run))
    if rank == 0:
        os.makedirs(base, exist_ok=True)
    if dist.is_available() and dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.barrier()
    path = os.path.join(base, "gpu_" + str(rank))
    os.makedirs(path, exist_ok=True)
    return path


def get_ckpt_filename(path, epoch):
    return os.path.join(path, 'ckpt_' + str(epoch) + '.pth')


def _each_quantizer(model):
    for name, module in model.named_modules():
        if isinstance(module, Q):
            yield name, module


def disable_input_quantization(model):
    for _, q in _each_quantizer(model):
        q.disable_input_quantization()


def enable_quantization(model):
    for name, q in _each_quantizer(model):
        q.enable_quantization(name)


def disable_quantization(model):
    for name, q in _each_quantizer(model):
        q.disable_quantization(name)


def get_model(args):
    kw = dict(aux_logits=False) if args.model == "inception_v3" else {}
    return models.__dict__[args.model](pretrained=True, **kw)

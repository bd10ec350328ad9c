"""Host-side mirror of the antquant API (no GPU): names, model surgery, state_dict layout, toggles,
8-bit promotion, checkpoint shim -- against what the unmodified reference produced (tests/golden/model_*.json)."""
import json
import os

import pytest

from host_util import run

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("flavor,mode", [("ant", "ant-int-pot-flint"), ("olive", "ant-int-flint")])
def test_quantize_model_layout_matches_reference(flavor, mode):
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "model_%s.json" % flavor)))
    res, _ = run(flavor, r'''
set_quantizer(mkargs(%r))
net = Net().eval()
q = quantize_model(net)
enable_quantization(q)
RESULT["keys"] = list(q.state_dict().keys())
RESULT["types"] = {n: type(m).__name__ for n, m in q.named_modules()}
RESULT["names"] = {n: m.name for n, m in q.named_modules() if isinstance(m, TensorQuantizer)}
RESULT["untouched"] = all(type(m).__name__ != "LinearQuantizer" for m in net.modules())
sd = q.state_dict()
RESULT["shapes"] = {k: list(v.shape) for k, v in sd.items()}
RESULT["signed"] = {n: bool(m.is_signed) for n, m in q.named_modules() if isinstance(m, TensorQuantizer)}
RESULT["perchannel"] = {n: bool(m.is_perchannel) for n, m in q.named_modules() if isinstance(m, TensorQuantizer)}
try:
    q(torch.randn(2, 3, 8, 8)); RESULT["cpu_forward"] = "ran"
except RuntimeError as e:
    RESULT["cpu_forward"] = str(e)
disable_quantization(q)
RESULT["disabled_equal"] = bool(torch.equal(q(torch.ones(2, 3, 8, 8)), net(torch.ones(2, 3, 8, 8))))
''' % mode)
    assert res["keys"] == ref["keys"]
    assert res["types"]["blocks"] == "Sequential" and res["types"]["features.0"] == "Conv2dQuantizer"
    assert res["types"]["head"] == "LinearQuantizer" and res["untouched"]
    assert res["names"]["head.quant_input"] == "head.quant_input"
    assert res["shapes"]["features.0.quant_weight.alpha"] == [8, 1] and res["shapes"]["head.quant_input.alpha"] == []
    assert res["shapes"]["head.quant_weight.quant_grid"] == [16]
    assert all(res["signed"][k] == (not k.endswith("quant_input")) for k in res["signed"])      # before calibration
    assert all(res["perchannel"][k] == k.endswith("quant_weight") for k in res["perchannel"])
    assert "no CPU" in res["cpu_forward"] or "CUDA" in res["cpu_forward"]
    assert res["disabled_equal"]


def test_star_imports_leak_like_the_reference():
    res, _ = run("ant", r'''
RESULT["names"] = sorted(n for n in ("logging", "os", "torch", "dist", "quant_args", "set_quantizer", "quantize_model",
    "enable_quantization", "disable_quantization", "disable_input_quantization", "set_first_last_layer",
    "set_8_bit_layer_n", "set_8_bit_layer_l", "load_ant_state_dict", "get_model", "get_ckpt_path", "get_ckpt_filename",
    "set_util_logging", "tag_info", "TensorQuantizer", "Conv2dQuantizer", "LinearQuantizer", "MultiheadAttentionQuantizer",
    "np", "nn", "copy", "models", "uuid", "logger") if n in globals())
import quant_modules as qm
RESULT["qm"] = sorted(n for n in ("QuantBase", "Quantizer", "TensorQuantizer", "Conv2dQuantizer", "LinearQuantizer", "F", "Tensor",
    "np", "dist", "quant_cuda", "logging", "AsymmetricQuantFunction", "linear_quantize", "linear_dequantize", "clamp",
    "asymmetric_linear_quantization_params") if hasattr(qm, n))
RESULT["tag"] = [tag_info(types.SimpleNamespace(tag="")), tag_info(types.SimpleNamespace(tag="x"))]
RESULT["ckpt"] = get_ckpt_filename("p", 3)
''')
    assert len(res["names"]) == 29, res["names"]
    assert len(res["qm"]) == 16, res["qm"]
    assert res["tag"] == ["", "_x"] and res["ckpt"] == os.path.join("p", "ckpt_3.pth")


def test_toggles_bit_promotion_and_checkpoint_shim():
    res, out = run("ant", r'''
set_quantizer(mkargs("ant-int-flint"))
q = quantize_model(Net())
enable_quantization(q)
qs = [m for m in q.modules() if isinstance(m, TensorQuantizer)]
disable_input_quantization(q)
RESULT["act_off"] = [m.is_enable_activation for m in qs]
for i, m in enumerate(qs):
    m.mse = torch.tensor(float(i % 7))
    m.has_inited_quant_para.data = torch.tensor(1.0)
set_8_bit_layer_n(q, 3)
RESULT["bits_n"] = [int(m.bit) for m in qs]
RESULT["reset"] = [float(m.has_inited_quant_para) for m in qs]
q2 = quantize_model(Net()); enable_quantization(q2)
set_8_bit_layer_l(q2, "0,2")
RESULT["bits_l"] = [int(m.bit) for m in q2.modules() if isinstance(m, TensorQuantizer)]
set_8_bit_layer_l(q2, "None")
set_first_last_layer(q2)
# checkpoint shim: eval builds 8-bit tables, the checkpoint holds 4-bit ones
args8 = mkargs("ant", wbit=8, abit=8); set_quantizer(args8)
q8 = quantize_model(Net())
set_quantizer(mkargs("ant-int-flint"))
ck = quantize_model(Net()).state_dict()
RESULT["before"] = list(q8.state_dict()["head.quant_weight.quant_grid"].shape)
load_ant_state_dict(q8, ck)
q8.load_state_dict(ck, strict=True)
RESULT["after"] = list(q8.state_dict()["head.quant_weight.quant_grid"].shape)
tq = qs[0]
RESULT["grids"] = {k: getattr(tq, k + "_value")().tolist() for k in ("int", "flint", "pot", "float", "apot")}
RESULT["float2"] = tq.float_value(2).tolist()
''')
    assert res["act_off"] == [False] * 10
    # the last two layers (4 quantizers) always, plus the pair with the largest summed mse among the rest
    assert res["bits_n"][-4:] == [8, 8, 8, 8] and res["bits_n"].count(8) == 6
    assert res["bits_n"][4:6] == [8, 8]                     # pairs: (0+1), (2+3), (4+5)=9 is the largest
    assert res["reset"] == [0.0] * 10
    assert res["bits_l"] == [8, 8, 4, 4, 8, 8, 4, 4, 4, 4]
    assert res["before"] == [256] and res["after"] == [16]
    assert "8-bit Re-SET" in out
    import numpy as np
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "grids.npz")))
    for k, key in (("int", "ant_int_s_4"), ("flint", "ant_flint_s_4"), ("pot", "ant_pot_s_4"), ("float", "ant_float3_s_4"),
                   ("apot", "ant_apot_s_4")):
        assert np.array_equal(np.array(res["grids"][k], dtype=np.float32), g[key]), k
    assert np.array_equal(np.array(res["float2"], dtype=np.float32), g["ant_float2_s_4"])


def test_olive_specifics():
    res, _ = run("olive", r'''
set_quantizer(mkargs("ant-int-flint", w_up=250, a_up=250))
class LM(nn.Module):
    def __init__(self):
        super().__init__()
        self.body = nn.Linear(8, 8)
        self.lm_head = nn.Linear(8, 16)
q = quantize_model(LM())
RESULT["types"] = {n: type(m).__name__ for n, m in q.named_children()}
tq = q.body.quant_weight
RESULT["bufs"] = sorted(k for k, _ in tq.named_buffers())
RESULT["int"] = tq.int_value().tolist(); RESULT["flint"] = tq.flint_value().tolist(); RESULT["out"] = tq.outlier_value().tolist()
try:
    from transformers import pytorch_utils
    c = pytorch_utils.Conv1D(12, 8)
    RESULT["conv1d"] = type(quantize_model(nn.Sequential(c))[0]).__name__
except ImportError:
    RESULT["conv1d"] = "Conv1dQuantizer"
import quant_modules as qm
RESULT["has_dist"] = hasattr(qm, "dist")
''')
    assert res["types"] == {"body": "LinearQuantizer", "lm_head": "Linear"}
    assert res["bufs"] == ["bit", "has_inited_quant_para", "outliers", "quant_grid"]
    assert res["int"] == [float(4 * i) for i in range(-7, 8)]
    assert res["flint"] == [-32, -16, -12, -8, -6, -4, -2, 0, 2, 4, 6, 8, 12, 16, 32]
    assert res["out"] == [-384, -256, -192, -128, -96, -64, -48, 48, 64, 96, 128, 192, 256, 384]
    assert res["conv1d"] == "Conv1dQuantizer"
    assert res["has_dist"] is False


@pytest.mark.parametrize("flavor,mode", [("ant", "ant-int-pot-flint"), ("olive", "ant-int-flint")])
def test_weight_cache_key_host_logic(flavor, mode):
    """The weight-quant cache key (antq/layers.py) without a GPU: absent before calibration and under QAT autograd,
    present in eval, and different after every change the quantized weight depends on."""
    res, _ = run(flavor, r'''
import antq.layers as L
lin = nn.Linear(64, 32)
q = LinearQuantizer(mode=%r, wbit=4, abit=4, args=mkargs(%r))
q.set_param(lin)
q.eval()
q.quant_weight.enable_quantization("w")
RESULT["before_init"] = q._weight_key() is None
q.quant_weight.has_inited_quant_para.data = torch.ones_like(q.quant_weight.has_inited_quant_para)
with torch.no_grad():
    k0 = q._weight_key()
    q.weight.mul_(2.0)
    k1 = q._weight_key()
    q.quant_weight.alpha.data = q.quant_weight.alpha.data * 0.5
    k2 = q._weight_key()
    q.quant_weight.quant_grid.data = q.quant_weight.quant_grid.data.clone()
    k3 = q._weight_key()
    q.quant_weight.is_enable_weight = False
    k4 = q._weight_key()
    q.quant_weight.is_enable_weight = True
RESULT["keys_differ"] = k0 is not None and len({k0, k1, k2, k3, k4}) == 5
RESULT["stable"] = q._weight_key() == q._weight_key()
RESULT["under_grad"] = q._weight_key() is None if %r == "ant" else q._weight_key() is not None
with torch.no_grad():
    q.train(); RESULT["train_off"] = q._weight_key() is None; q.eval()
L.CACHE_WEIGHTS = False
with torch.no_grad():
    RESULT["switched_off"] = q._weight_key() is None
''' % (mode, mode, flavor))
    assert res["before_init"] and res["keys_differ"] and res["stable"] and res["under_grad"] and res["switched_off"] and res["train_off"], res

"""Random-init architectures + synthetic inputs for the model-shaped parity configs (BASELINE.json C2-C4,
SURVEY.md section 8(d)): no weights or datasets exist offline, so every model is built from a fixed CPU seed.
The SAME function runs in the build container (where tests/golden/make_golden_models.py pushes the model through
the UNMODIFIED reference `quantize_model`) and on the GPU box (where the tests push it through this repo's
mirror), so both sides see bit-identical fp32 weights and inputs.

Reference hooks these stand for:  A/ImageNet/main.py:117-128 (ResNet-50, ViT), A/BERT/run_glue.py:538-546 (BERT),
O/llm/run_clm.py:603-613 (OPT, GPT-2).
"""
import types

import torch


def _args(mode, **kw):
    d = dict(mode=mode, wbit=4, abit=4, w_up=150, a_up=150, w_low=75, a_low=75, percent=100, search=False,
             no_outlier=False)
    d.update(kw)
    return types.SimpleNamespace(**d)


def resnet50():
    import torchvision
    torch.manual_seed(0)
    m = torchvision.models.resnet50(weights=None).eval()
    x = torch.randn(1, 3, 64, 64)
    return m, (x,), {}, _args("ant-int-pot-flint")


class _BertLayer(torch.nn.Module):
    def __init__(self, h, heads, ffn):
        super().__init__()
        nn = torch.nn
        self.heads = heads
        self.query, self.key, self.value = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
        self.attn_out, self.attn_ln = nn.Linear(h, h), nn.LayerNorm(h, eps=1e-12)
        self.inter, self.out, self.out_ln = nn.Linear(h, ffn), nn.Linear(ffn, h), nn.LayerNorm(h, eps=1e-12)

    def forward(self, x):
        B, S, H = x.shape
        sp = lambda t: t.view(B, S, self.heads, H // self.heads).permute(0, 2, 1, 3)
        q, k, v = sp(self.query(x)), sp(self.key(x)), sp(self.value(x))
        p = torch.softmax(q @ k.transpose(-1, -2) / (H // self.heads) ** 0.5, dim=-1)
        ctx = (p @ v).permute(0, 2, 1, 3).reshape(B, S, H)
        x = self.attn_ln(self.attn_out(ctx) + x)
        return self.out_ln(self.out(torch.nn.functional.gelu(self.inter(x))) + x)


class BertLike(torch.nn.Module):
    """BERT-base shaped encoder (A/BERT/bert_config.json: hidden 768, 12 heads, ffn 3072) with the module layout of
    A/BERT/modeling.py (nn.Linear query / key / value / dense, pooler, classifier), cut to 2 layers.  HF's BertModel
    cannot go through the ANT tree's quantize_model: its `base_model` property recurses forever
    (A/antquant/quant_model.py:45-49 has no skip list, unlike O/antquant/quant_model.py:50)."""

    def __init__(self, vocab=30522, h=768, heads=12, ffn=3072, layers=2, labels=3, max_pos=512):
        super().__init__()
        nn = torch.nn
        self.word, self.pos, self.typ = nn.Embedding(vocab, h), nn.Embedding(max_pos, h), nn.Embedding(2, h)
        self.emb_ln = nn.LayerNorm(h, eps=1e-12)
        self.layer = nn.ModuleList([_BertLayer(h, heads, ffn) for _ in range(layers)])
        self.pooler, self.classifier = nn.Linear(h, h), nn.Linear(h, labels)

    def forward(self, ids):
        pos = torch.arange(ids.shape[1], device=ids.device).unsqueeze(0)
        x = self.emb_ln(self.word(ids) + self.pos(pos) + self.typ(torch.zeros_like(ids)))
        for l in self.layer:
            x = l(x)
        return self.classifier(torch.tanh(self.pooler(x[:, 0])))


def bert2():
    torch.manual_seed(1)
    m = BertLike().eval()
    for p in m.parameters():                       # BERT's initializer_range
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.02)
    ids = torch.randint(0, 30522, (2, 64))
    return m, (ids,), {}, _args("ant-int-pot-float")


def opt2():
    from transformers import OPTConfig, OPTForCausalLM
    torch.manual_seed(2)
    cfg = OPTConfig(vocab_size=1024, hidden_size=512, ffn_dim=2048, num_hidden_layers=2, num_attention_heads=8,
                    max_position_embeddings=128, word_embed_proj_dim=512, attn_implementation="eager")
    m = OPTForCausalLM(cfg).eval()
    ids = torch.randint(0, 1024, (2, 64))
    return m, (ids,), {}, _args("ant-int-flint", w_up=250, a_up=250)


def gpt2():
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(3)
    cfg = GPT2Config(vocab_size=512, n_positions=64, n_embd=128, n_layer=2, n_head=4, attn_implementation="eager")
    m = GPT2LMHeadModel(cfg).eval()
    ids = torch.randint(0, 512, (2, 48))
    return m, (ids,), {}, _args("ant-int-flint", w_up=250, a_up=250)


def vit():
    import torchvision
    torch.manual_seed(4)
    m = torchvision.models.VisionTransformer(image_size=32, patch_size=8, num_layers=2, num_heads=4, hidden_dim=64,
                                             mlp_dim=128, num_classes=10).eval()
    torch.nn.init.normal_(m.heads.head.weight, std=0.05)      # torchvision zero-inits the head: alpha = 0, NaN logits
    x = torch.randn(2, 3, 32, 32)
    return m, (x,), {}, _args("ant-int-pot-flint", w_low=80, a_low=40)


ZOO = {"resnet50": ("ant", resnet50), "bert2": ("ant", bert2), "vit": ("ant", vit),
       "opt2": ("olive", opt2), "gpt2": ("olive", gpt2)}


def logits_of(out):
    if isinstance(out, torch.Tensor):
        return out
    if hasattr(out, "logits"):
        return out.logits
    return out[0]


def fp32_checksum(model):
    """Guards the 'same seed -> same weights' assumption across boxes."""
    s = 0.0
    for _, p in sorted(model.state_dict().items()):
        if p.dtype.is_floating_point:
            s += float(p.double().abs().sum())
    return s

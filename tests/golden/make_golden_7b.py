"""Generates tests/golden/vidi7b_reference_golden.pt by running the REFERENCE'S OWN Vidi-7B decoder layer
(Vidi_7B/model/lmm/dattn/mistral.py: DattnMistralDecoderLayer.forward, DattnMistralFlashAttention2.forward_xattn;
xattn.py: flash_cross_attention_forward; split.py: splitted_call), imported unmodified from /root/reference.

Third-party pieces it cannot get here and how they are supplied:
  * transformers==4.44.2's ``MistralFlashAttention2`` (the base class of the reference's attention) no longer exists in
    the installed transformers 5.5.0 -> a stand-in with the same parameters whose ``forward`` restates its documented
    semantics (q/k/v proj, rotate-half RoPE theta=config.rope_theta, causal softmax with GQA, o_proj) in fp32;
  * ``flash_attn_func`` -> eager fp32 restatement (ref_shim.eager_flash_attn_func);
  * MistralRMSNorm / MistralMLP / MistralConfig are the installed HF classes.
Run in the build container: ``python tests/golden/make_golden_7b.py``.
"""
import importlib.machinery
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


for n in ("langid", "orjson"):
    _stub(n)
_stub("decord", VideoReader=object, cpu=lambda *a, **k: None)
import transformers.utils as tu  # noqa: E402

tu.is_flash_attn_2_available = lambda: True
tu.is_flash_attn_greater_or_equal = lambda v: True
import transformers.models.mistral.modeling_mistral as mm  # noqa: E402


class MistralFlashAttention2(nn.Module):
    """Stand-in for transformers 4.44.2 MistralFlashAttention2 (parameters + T2T forward semantics)."""

    def __init__(self, config, layer_idx=None):
        super().__init__()
        self.config, self.layer_idx = config, layer_idx
        self.hidden_size, self.num_heads = config.hidden_size, config.num_attention_heads
        self.head_dim = getattr(config, "head_dim", None) or self.hidden_size // self.num_heads
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.attention_dropout = 0.0
        self.q_proj = nn.Linear(self.hidden_size, self.num_heads * self.head_dim, bias=False)
        self.k_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=False)
        self.v_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=False)
        self.o_proj = nn.Linear(self.num_heads * self.head_dim, self.hidden_size, bias=False)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None, output_attentions=False,
                use_cache=False, cache_position=None, **kw):
        B, T, _ = hidden_states.shape
        H, Hk, dh = self.num_heads, self.num_key_value_heads, self.head_dim
        q = self.q_proj(hidden_states).view(B, T, H, dh).transpose(1, 2)
        k = self.k_proj(hidden_states).view(B, T, Hk, dh).transpose(1, 2)
        v = self.v_proj(hidden_states).view(B, T, Hk, dh).transpose(1, 2)
        inv = 1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.float) / dh))
        fr = position_ids[0].float()[:, None] * inv[None]
        emb = torch.cat([fr, fr], -1)
        cos, sin = emb.cos()[None, None], emb.sin()[None, None]
        rot = lambda x: torch.cat([-x[..., dh // 2:], x[..., :dh // 2]], -1)
        q, k = q * cos + rot(q) * sin, k * cos + rot(k) * sin
        k, v = k.repeat_interleave(H // Hk, 1), v.repeat_interleave(H // Hk, 1)
        s = (q @ k.transpose(-1, -2)) * dh ** -0.5
        s = s + torch.full((T, T), float("-inf")).triu(1)
        o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, H * dh)
        return self.o_proj(o), None, past_key_value


mm.MistralFlashAttention2 = MistralFlashAttention2
sys.path.insert(0, "/root/reference/Vidi_7B")
import model.lmm.dattn.mistral as M  # noqa: E402  (reference code, unmodified)
import model.lmm.dattn.xattn as X7  # noqa: E402

X7.flash_attn_func = ref_shim.eager_flash_attn_func
X7.flash_attn_varlen_func = ref_shim.eager_flash_attn_varlen_func
_unpad_installed = X7.unpad_input                      # flash-attn 2.8.3 returns 5 values, the pinned 2.6.3 returned 4
X7.unpad_input = lambda h, m: _unpad_installed(h, m)[:4]

from vidi_b200 import synth  # noqa: E402
from vidi_b200.config import AudioCfg, MistralCfg, Vidi7BConfig, VisionCfg  # noqa: E402

torch.manual_seed(7)
cfg = Vidi7BConfig(llm=MistralCfg(hidden=64, heads=8, kv_heads=2, head_dim=8, inter=128, layers=2, vocab=128),
                   vis=VisionCfg(hidden=32, heads=2, inter=48, layers=3, image=378, patch=14),
                   aud=AudioCfg(d_model=32, heads=2, ffn=64, layers=2), mm_image_pool_size=4, name="golden-tiny-7b")
sd = synth.make_state_dict(cfg, seed=778)
mcfg = mm.MistralConfig(hidden_size=64, num_attention_heads=8, num_key_value_heads=2, intermediate_size=128, num_hidden_layers=2,
                        vocab_size=128, rms_norm_eps=1e-5, hidden_act="silu", head_dim=8)
mcfg.mm_splits = 2
mcfg._attn_implementation = "eager"
T, Ni, Na = 7, 40, 23
H = torch.randn(1, T, 64)
img = torch.randn(1, Ni, 64) * 0.5
aud = torch.randn(1, Na, 64) * 0.5
pos_ids = torch.arange(T)[None]
m_img, m_aud = torch.ones(1, Ni, dtype=torch.bool), torch.ones(1, Na, dtype=torch.bool)
out_layers = []
hs, im, au = H, img, aud
for l in range(2):
    layer = M.DattnMistralDecoderLayer(mcfg, l).eval()
    layer.load_state_dict({k[len(f"model.layers.{l}."):]: v for k, v in sd.items() if k.startswith(f"model.layers.{l}.")})
    with torch.no_grad():
        (hs_out,), im_out, au_out = layer(hs, attention_mask=torch.ones(1, T, dtype=torch.long), position_ids=pos_ids,
                                          image_embeds=im, image_attention_mask=m_img, audio_embeds=au,
                                          audio_attention_mask=m_aud, past_key_value=None, past_image_key_value=None,
                                          past_audio_key_value=None, use_cache=False, cache_position=torch.arange(T))
    out_layers.append(dict(text=hs_out[0], image=im_out[0], audio=au_out[0]))
    hs, im, au = hs_out, im_out, au_out
OUT = dict(seed=778, cfg=dict(llm=vars(cfg.llm), vis=vars(cfg.vis), aud=vars(cfg.aud), pool=cfg.mm_image_pool_size),
           H0=H[0], img0=img[0], aud0=aud[0], layers=out_layers)
torch.save(OUT, os.path.join(HERE, "vidi7b_reference_golden.pt"))
print("wrote vidi7b_reference_golden.pt", os.path.getsize(os.path.join(HERE, "vidi7b_reference_golden.pt")), "bytes")

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

# ---------------------------------------------------------------- whole model: encoders + input prep + layer loop + lm_head
# DattnMistralForCausalLM.forward (mistral.py:512-616) -> DattnMMMixin.prepare_inputs_labels_for_multimodal / encode_video_images /
# encode_video_audios (Vidi_7B/model/lmm/dattn/multimodal.py:154-227) -> DattnMistralModel.forward (mistral.py:296-453), all run
# UNMODIFIED on instances assembled around the reference layers above and the reference's own leaf modules (the constructors
# want hub downloads and flash-attn; forward() does not).  Towers: installed HF SiglipVisionModel / WhisperEncoder.
from types import SimpleNamespace  # noqa: E402

from transformers import SiglipVisionConfig, SiglipVisionModel, WhisperConfig  # noqa: E402
from transformers.models.whisper.modeling_whisper import WhisperEncoder  # noqa: E402
from model.mm_layer import MLP, RMSNorm  # noqa: E402  (reference code)
from model.mm_vision import Conv2DPool, LearnablePosEmbd  # noqa: E402  (reference code)

D = cfg.llm.hidden
vcfg = SiglipVisionConfig(hidden_size=32, intermediate_size=48, num_hidden_layers=3, num_attention_heads=2, image_size=378,
                          patch_size=14, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)
vcfg._attn_implementation = "eager"
vis = SiglipVisionModel(vcfg).eval()
_miss = vis.load_state_dict({k[len("model.mm_vis."):]: v for k, v in sd.items() if k.startswith("model.mm_vis.")}, strict=False)
assert all("head" in k for k in _miss.missing_keys), _miss
acfg = WhisperConfig(d_model=32, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=64, num_mel_bins=128,
                     max_source_positions=1500, activation_function="gelu")
acfg._attn_implementation = "eager"
aud_enc = WhisperEncoder(acfg).eval()
aud_enc.load_state_dict({k[len("model.mm_aud.encoder."):]: v for k, v in sd.items() if k.startswith("model.mm_aud.encoder.")})


class VisTower(nn.Module):
    """Same contract as SiglipVisionTower.forward (Vidi_7B/model/mm_vision/siglip.py) around the HF model built above."""
    num_patches_per_side = 27
    hidden_size = 32

    def __init__(self, m):
        super().__init__()
        self.vision_model = m.vision_model
        self.m = m

    def forward(self, imgs):
        o = self.m(imgs, output_hidden_states=True)
        return o.pooler_output, o.hidden_states[-2]


class AudTower(nn.Module):
    hidden_size = 32

    def __init__(self, enc):
        super().__init__()
        self.encoder = enc
        self.config = enc.config

    def forward(self, a):
        return self.encoder(a)[0]


def load_mod(mod, prefix):
    mod.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)})
    return mod.eval()


mdl = M.DattnMistralModel.__new__(M.DattnMistralModel)
nn.Module.__init__(mdl)
mdl.config = mcfg
mdl.embed_tokens = nn.Embedding(cfg.llm.vocab, D)
mdl.embed_tokens.weight.data.copy_(sd["model.embed_tokens.weight"])
ref_layers = []
for l in range(2):
    layer = M.DattnMistralDecoderLayer(mcfg, l).eval()
    layer.load_state_dict({k[len(f"model.layers.{l}."):]: v for k, v in sd.items() if k.startswith(f"model.layers.{l}.")})
    ref_layers.append(layer)
mdl.layers = nn.ModuleList(ref_layers)
mdl.norm = load_mod(mm.MistralRMSNorm(D, eps=1e-5), "model.norm.")
mdl.gradient_checkpointing = False
pool_s = cfg.mm_image_pool_size
mdl.mm_vis, mdl.mm_aud = VisTower(vis), AudTower(aud_enc)
mdl.audio_processor = SimpleNamespace(nb_max_frames=3000)
mdl.text_tokenizer = SimpleNamespace(padding_side="right")
mdl.mm_rand_img_pool = load_mod(Conv2DPool(d_in=32, d_out=32, s_in=27, s_out=pool_s), "model.mm_rand_img_pool.")
mdl.mm_rand_img_projector = load_mod(MLP("mlp2x_gelu", 32, D), "model.mm_rand_img_projector.")
mdl.mm_rand_img_norm = load_mod(RMSNorm(D), "model.mm_rand_img_norm.")
mdl.mm_rand_aud_pool = load_mod(nn.Conv1d(32, 32, cfg.mm_audio_pool_size, stride=cfg.mm_audio_pool_size, bias=False), "model.mm_rand_aud_pool.")
mdl.mm_rand_aud_projector = load_mod(MLP("mlp2x_gelu", 32, D), "model.mm_rand_aud_projector.")
mdl.mm_rand_aud_norm = load_mod(RMSNorm(D), "model.mm_rand_aud_norm.")
mdl.mm_rand_llm_norm = load_mod(RMSNorm(D), "model.mm_rand_llm_norm.")
mdl.mm_rand_pos_h = load_mod(LearnablePosEmbd(D, pool_s), "model.mm_rand_pos_h.")
mdl.mm_rand_pos_w = load_mod(LearnablePosEmbd(D, pool_s), "model.mm_rand_pos_w.")
mdl.mm_rand_pos_t = load_mod(LearnablePosEmbd(D, cfg.mm_time_interval), "model.mm_rand_pos_t.")
top = M.DattnMistralForCausalLM.__new__(M.DattnMistralForCausalLM)
nn.Module.__init__(top)
top.model = mdl.eval()
top.lm_head = nn.Linear(D, cfg.llm.vocab, bias=False)
top.lm_head.weight.data.copy_(sd["lm_head.weight"])
for _k, _v in dict(train_vis=False, train_aud=False, mm_splits=2, mm_image_pool_size=pool_s, mm_audio_pool_size=cfg.mm_audio_pool_size,
                   mm_input_type="video", loss_thres=None).items():
    setattr(mcfg, _k, _v)
top.config = mcfg
top.vocab_size = cfg.llm.vocab
top.eval()
ids, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=6, seed=98, audio_size=1234)
with torch.no_grad():
    X_img, m_img2 = top.encode_video_images([images])
    X_aud, m_aud2 = top.encode_video_audios([mels], [asz])
    co = M.DattnMistralForCausalLM.forward(
        top, input_ids=ids[None] if ids.dim() == 1 else ids, attention_mask=torch.ones(1, ids.numel(), dtype=torch.long),
        images=[images], audios=[mels], audio_sizes=[asz], use_cache=False,
        output_attentions=False, output_hidden_states=False, return_dict=True)
OUT["model"] = dict(inputs="synth.make_inputs(cfg, 3, 1, n_text=6, seed=98, audio_size=1234)", audio_size=asz,
                    image_embeds=X_img[0], image_mask=m_img2[0], audio_embeds=X_aud[0], audio_mask=m_aud2[0],
                    logits=co.logits[0].float())
torch.save(OUT, os.path.join(HERE, "vidi7b_reference_golden.pt"))
print("wrote vidi7b_reference_golden.pt", os.path.getsize(os.path.join(HERE, "vidi7b_reference_golden.pt")), "bytes")

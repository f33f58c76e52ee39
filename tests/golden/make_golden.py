"""Generates tests/golden/*.pt by RUNNING THE REFERENCE'S OWN CODE (imported from /root/reference through
ref_shim.py) and the third-party HF blocks it subclasses, on seeded tiny inputs.  Run in the build container:

    python tests/golden/make_golden.py

The fixtures pin the oracle (tests/test_oracle_golden.py) and travel to the GPU box; /root/reference does not.

What runs as reference code, unmodified:
  * vidi/model/mm_vision/pool.py  Conv2DPool.forward          * vidi/utils.py space_to_depth, resize_by_tokens
  * vidi/model/mm_vision/pos.py   LearnablePosEmbd.forward     * vidi/model/mm_layer/{norm,mlp}.py RMSNorm, rms_norm, MLP
  * vidi/model/lmm/dattn/multimodal.py  DattnMMMixin.encode_video_images / encode_video_audios   (on HF tower modules)
  * vidi/model/lmm/dattn/gemma.py       DattnGemma2DecoderLayer.forward, DattnGemma2Attention.forward_xattn
  * vidi/model/lmm/dattn/xattn.py       flash_cross_attention_forward      * vidi/model/lmm/dattn/split.py splitted_call
Third-party pieces: HF transformers (installed 5.5.0; the reference pins 4.50.0) SiglipVisionModel, WhisperEncoder,
Gemma2Attention/RMSNorm/MLP/RotaryEmbedding run as shipped; flash_attn_func is replaced by an eager fp32 restatement
(ref_shim.eager_flash_attn_func) because FA2 has no CPU kernel.
"""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

G, X = ref_shim.install()
from vidi.model.mm_layer import MLP, RMSNorm, rms_norm  # noqa: E402
from vidi.model.mm_vision.pool import Conv2DPool  # noqa: E402
from vidi.model.mm_vision.pos import LearnablePosEmbd  # noqa: E402
from vidi.utils import resize_by_tokens, space_to_depth  # noqa: E402

torch.manual_seed(20250922)
OUT = {}


def rnd(*s, scale=1.0):
    return torch.randn(*s) * scale


# ---------------------------------------------------------------- leaf modules
with torch.no_grad():
    x = rnd(2, 6, 27, 27)
    pool = Conv2DPool(6, 6, 27, 2, mm_splits=1, mm_image_pool_size=2)
    OUT["pool"] = dict(x=x, out={f"{h}x{w}": pool(x, (h, w)) for h, w in [(28, 28), (20, 20), (10, 10), (16, 12)]})
    OUT["s2d"] = dict(x=rnd(1, 3, 4, 6), out=space_to_depth(OUT.get("s2d_x", rnd(1, 3, 4, 6)), 2))
    xs = rnd(1, 3, 4, 6)
    OUT["s2d"] = dict(x=xs, out=space_to_depth(xs, 2))
    OUT["resize_by_tokens"] = {B: resize_by_tokens(torch.zeros(B, 1, 27, 27), 240000) for B in (307, 400, 600, 1200, 1800, 3600, 7200, 20000)}
    d = 32
    pe_h = LearnablePosEmbd(d, 2).eval()
    pe_t = LearnablePosEmbd(d, 10000).eval()
    OUT["pos"] = dict(
        h_sd={k: v.clone() for k, v in pe_h.state_dict().items()}, t_sd={k: v.clone() for k, v in pe_t.state_dict().items()},
        h_out=pe_h(torch.zeros(3, 14, 14, d), dim=1).reshape(14, d), t_out=pe_t(torch.zeros(37, 2, 2, d), dim=0).reshape(37, d))
    nrm = RMSNorm(d, std=0.7); nrm.weight.data += rnd(d, scale=0.1)
    xn = rnd(5, d, scale=3.0)
    OUT["norm"] = dict(x=xn, w=nrm.weight.data.clone(), out=nrm(xn), out_plain=rms_norm(xn))
    mlp = MLP("mlp2x_gelu", 24, d).eval()
    xm = rnd(7, 24)
    OUT["mlp"] = dict(x=xm, sd={k: v.clone() for k, v in mlp.state_dict().items()}, out=mlp(xm))

# ---------------------------------------------------------------- tiny Vidi1.5 config shared by the rest
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from vidi_b200.config import AudioCfg, LLMCfg, Vidi15Config, VisionCfg  # noqa: E402
from vidi_b200 import synth  # noqa: E402

cfg = Vidi15Config(llm=LLMCfg(hidden=64, heads=4, kv_heads=2, head_dim=16, inter=128, layers=2, vocab=128, query_pre_attn_scalar=16.0),
                   vis=VisionCfg(hidden=32, heads=2, inter=48, layers=3, image=378, patch=14),
                   aud=AudioCfg(d_model=32, heads=2, ffn=64, layers=2), name="golden-tiny")
sd = synth.make_state_dict(cfg, seed=777)
OUT["cfg"] = dict(llm=vars(cfg.llm), vis=vars(cfg.vis), aud=vars(cfg.aud))
OUT["seed"] = 777

from transformers import SiglipVisionConfig, SiglipVisionModel, WhisperConfig  # noqa: E402
from transformers.models.whisper.modeling_whisper import WhisperEncoder  # noqa: E402

vcfg = SiglipVisionConfig(hidden_size=32, intermediate_size=48, num_hidden_layers=3, num_attention_heads=2, image_size=378,
                          patch_size=14, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)
vcfg._attn_implementation = "eager"
vis = SiglipVisionModel(vcfg).eval()
missing = vis.load_state_dict({k[len("model.mm_vis."):]: v for k, v in sd.items() if k.startswith("model.mm_vis.")}, strict=False)
assert all("head" in k for k in missing.missing_keys), missing
acfg = WhisperConfig(d_model=32, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=64, num_mel_bins=128,
                     max_source_positions=1500, activation_function="gelu")
acfg._attn_implementation = "eager"
aud = WhisperEncoder(acfg).eval()
aud.load_state_dict({k[len("model.mm_aud.encoder."):]: v for k, v in sd.items() if k.startswith("model.mm_aud.encoder.")})

ids, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=6, seed=99, audio_size=1234)
with torch.no_grad():
    vo = vis(images, output_hidden_states=True)
    OUT["siglip"] = dict(hidden_m2=vo.hidden_states[-2])
    OUT["whisper"] = dict(out=aud(mels)[0])


# ---------------------------------------------------------------- reference encode_video_images / encode_video_audios
class VisTower(torch.nn.Module):
    """Same contract as SiglipVisionTower.forward (siglip.py:29-34) around the HF model built above."""
    num_patches_per_side = 27
    hidden_size = 32

    def __init__(self, m):
        super().__init__()
        self.vision_model = m.vision_model
        self.m = m

    def forward(self, imgs):
        o = self.m(imgs, output_hidden_states=True)
        return o.pooler_output, o.hidden_states[-2]


class AudTower(torch.nn.Module):
    def __init__(self, enc):
        super().__init__()
        self.encoder = enc
        self.config = enc.config

    def forward(self, a):
        return self.encoder(a)[0]


def load_mod(mod, prefix):
    mod.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)})
    return mod.eval()


D = cfg.llm.hidden
inner = SimpleNamespace(
    mm_vis=VisTower(vis), mm_aud=AudTower(aud), audio_processor=SimpleNamespace(nb_max_frames=3000),
    mm_rand_img_pool=Conv2DPool(32, 32, 27, 2, mm_splits=2, mm_image_pool_size=2),
    mm_rand_img_projector=load_mod(MLP("mlp2x_gelu", 32 * 4, D), "model.mm_rand_img_projector."),
    mm_rand_img_norm=load_mod(RMSNorm(D), "model.mm_rand_img_norm."),
    mm_rand_aud_pool=load_mod(torch.nn.Conv1d(32, D, 5, stride=5, bias=False), "model.mm_rand_aud_pool."),
    mm_rand_aud_projector=load_mod(MLP("mlp2x_gelu", D, D), "model.mm_rand_aud_projector."),
    mm_rand_aud_norm=load_mod(RMSNorm(D), "model.mm_rand_aud_norm."),
    mm_rand_llm_norm=load_mod(RMSNorm(D), "model.mm_rand_llm_norm."),
    mm_rand_pos_h=load_mod(LearnablePosEmbd(D, 2), "model.mm_rand_pos_h."),
    mm_rand_pos_w=load_mod(LearnablePosEmbd(D, 2), "model.mm_rand_pos_w."),
    mm_rand_pos_t=load_mod(LearnablePosEmbd(D, 10000), "model.mm_rand_pos_t."),
)


class Host(G.DattnMMMixin):
    training = False
    config = SimpleNamespace(train_vis=False, train_aud=False, mm_splits=2, mm_image_pool_size=2, mm_audio_pool_size=5)

    def get_model(self):
        return inner

    def encode_images(self, *a):
        raise NotImplementedError

    def prepare_inputs_labels_for_multimodal(self, *a):
        raise NotImplementedError


host = Host()
with torch.no_grad():
    X_img, m_img = host.encode_video_images([images])
    X_aud, m_aud = host.encode_video_audios([mels], [asz])
OUT["encode"] = dict(inputs="synth.make_inputs(cfg, 3, 1, n_text=6, seed=99, audio_size=1234)", audio_size=asz, image_embeds=X_img[0], image_mask=m_img[0],
                     audio_embeds=X_aud[0], audio_mask=m_aud[0])

# ---------------------------------------------------------------- reference decoder layer
gcfg = G.DattnGemma2Config(hidden_size=64, num_attention_heads=4, num_key_value_heads=2, head_dim=16, intermediate_size=128,
                           num_hidden_layers=2, vocab_size=128, query_pre_attn_scalar=16, sliding_window=4096,
                           attn_logit_softcapping=50.0, final_logit_softcapping=30.0, rms_norm_eps=1e-6,
                           hidden_activation="gelu_pytorch_tanh")
gcfg._attn_implementation = "eager"
gcfg.mm_splits = 2
from transformers.models.gemma2.modeling_gemma2 import Gemma2RotaryEmbedding  # noqa: E402

layer_out = []
ref_layers = []
T = 6
Himg, Haud = X_img.shape[1], X_aud.shape[1]
ids_noimg = ids[ids != -200]
nrmz = float(torch.tensor(D ** 0.5))
H0 = sd["model.embed_tokens.weight"][ids_noimg][None] * nrmz
img0, aud0 = X_img * nrmz, X_aud * nrmz
rot = Gemma2RotaryEmbedding(config=gcfg)
pos_ids = torch.arange(T)[None]
cos, sin = rot(H0, pos_ids)
hs, im, au = H0, img0, aud0
for l in range(2):
    layer = G.DattnGemma2DecoderLayer(gcfg, l).eval()
    layer.load_state_dict({k[len(f"model.layers.{l}."):]: v for k, v in sd.items() if k.startswith(f"model.layers.{l}.")})
    hf_forward = layer.self_attn.forward

    def t2t(hidden_states, position_embeddings, attention_mask=None, _f=hf_forward, **kw):
        # the reference hands FA2 a 2-D padding mask and FA2 applies causality itself; the eager HF path wants the 4-D additive mask
        Tq = hidden_states.shape[1]
        m4 = torch.full((Tq, Tq), float("-inf")).triu(1)[None, None]
        return _f(hidden_states, position_embeddings, m4, **{k: v for k, v in kw.items() if k in ("cache_position",)})
    layer.self_attn.forward = t2t
    with torch.no_grad():
        (hs_out,), im_out, au_out = layer(
            hs, position_embeddings=(cos, sin), attention_mask=torch.ones(1, T, dtype=torch.long), position_ids=pos_ids,
            image_embeds=im, image_attention_mask=m_img, audio_embeds=au, audio_attention_mask=m_aud,
            past_key_value=None, past_image_key_value=None, past_audio_key_value=None, use_cache=False,
            cache_position=torch.arange(T))
    layer_out.append(dict(text=hs_out[0], image=im_out[0], audio=au_out[0]))
    ref_layers.append(layer)
    hs, im, au = hs_out, im_out, au_out
OUT["decoder"] = dict(ids=ids, H0=H0[0], img0=img0[0], aud0=aud0[0], layers=layer_out)

# ---------------------------------------------------------------- Vidi-7B learned pool (Vidi_7B/model/mm_vision/pool.py)
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("vidi7b_pool", "/root/reference/Vidi_7B/model/mm_vision/pool.py")
_pool7 = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_pool7)
with torch.no_grad():
    x7 = rnd(2, 6, 27, 27)
    out7 = {}
    for s_out in (16, 4, 9):
        m7 = _pool7.Conv2DPool(6, 6, 27, s_out).eval()
        out7[s_out] = dict(w=m7.conv.weight.data.clone(), out=m7(x7))
    OUT["pool7b"] = dict(x=x7, cases=out7)

# ---------------------------------------------------------------- reference model-level loop, input prep and LM head
# DattnGemma2Model.forward (gemma.py:267-424), DattnGemma2ForCausalLM.forward (gemma.py:484-601) and
# DattnMMMixin.prepare_inputs_labels_for_multimodal (multimodal.py:339-451) run UNMODIFIED on instances assembled around the
# reference layers above (the constructors want flash-attn-2 and hub downloads; forward() itself does not).  use_cache=False: the
# HF 4.50 cache classes the cached branch builds no longer exist in 5.x.
from transformers.models.gemma2.modeling_gemma2 import Gemma2RMSNorm  # noqa: E402

mdl = G.DattnGemma2Model.__new__(G.DattnGemma2Model)
torch.nn.Module.__init__(mdl)
mdl.config = gcfg
mdl.embed_tokens = torch.nn.Embedding(cfg.llm.vocab, D)
mdl.embed_tokens.weight.data.copy_(sd["model.embed_tokens.weight"])
mdl.layers = torch.nn.ModuleList(ref_layers)
mdl.norm = Gemma2RMSNorm(D, eps=1e-6)
mdl.norm.weight.data.copy_(sd["model.norm.weight"])
mdl.rotary_emb = rot
mdl.gradient_checkpointing = False
for _k, _v in vars(inner).items():
    setattr(mdl, _k, _v)
mdl.text_tokenizer = SimpleNamespace(padding_side="right")
top = G.DattnGemma2ForCausalLM.__new__(G.DattnGemma2ForCausalLM)
torch.nn.Module.__init__(top)
top.model = mdl.eval()
top.lm_head = torch.nn.Linear(D, cfg.llm.vocab, bias=False)
top.lm_head.weight.data.copy_(sd.get("lm_head.weight", sd["model.embed_tokens.weight"]))
for _k, _v in dict(train_vis=False, train_aud=False, mm_splits=2, mm_image_pool_size=2, mm_audio_pool_size=5, mm_input_type="video").items():
    setattr(gcfg, _k, _v)
top.config = gcfg
top.vocab_size = cfg.llm.vocab
top.eval()
with torch.no_grad():
    mo = G.DattnGemma2Model.forward(
        mdl, inputs_embeds=sd["model.embed_tokens.weight"][ids_noimg][None], attention_mask=torch.ones(1, T, dtype=torch.long),
        image_embeds=X_img, image_attention_mask=m_img, audio_embeds=X_aud, audio_attention_mask=m_aud, use_cache=False,
        output_attentions=False, output_hidden_states=False, return_dict=True)
    co = G.DattnGemma2ForCausalLM.forward(
        top, input_ids=ids[None] if ids.dim() == 1 else ids, attention_mask=torch.ones(1, ids.numel(), dtype=torch.long),
        images=[images], audios=[mels], audio_sizes=[asz], use_cache=False,
        output_attentions=False, output_hidden_states=False, return_dict=True)
OUT["model"] = dict(last_hidden_state=mo.last_hidden_state[0], logits=co.logits[0].float())

torch.save(OUT, os.path.join(HERE, "vidi15_reference_golden.pt"))
sz = os.path.getsize(os.path.join(HERE, "vidi15_reference_golden.pt"))
print(f"wrote vidi15_reference_golden.pt ({sz / 1e6:.2f} MB); keys: {list(OUT)}")

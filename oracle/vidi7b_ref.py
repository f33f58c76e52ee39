"""ORACLE -- test infrastructure only.  fp32 CPU restatement of the Vidi-7B (Mistral Dattn) prefill.

Same rules as vidi15_ref.py (only tests / smoke / bench cpu legs may import it).  Follows
Vidi_7B/model/lmm/dattn/mistral.py:44-116 (forward_xattn), :131-137 (feed_foward), :139-274 (decoder layer),
:296-453 (model loop), :596-616 (lm_head, fp32 logits), Vidi_7B/model/lmm/dattn/multimodal.py:154-227 (encoders)
and Vidi_7B/model/mm_vision/pool.py:6-26 (learned conv pool + align_corners bilinear).
PARITY PINNING (tests/test_oracle_golden.py), all against the reference's own code run through tests/golden/make_golden_7b.py
(the transformers==4.44.2 MistralFlashAttention2 base class, gone from the installed 5.5.0, and flash-attn's kernels are
replaced there by fp32 restatements of their semantics): the learned-conv pool vs the reference's Conv2DPool; ``stream_layer``
+ ``text_layer`` vs two stacked DattnMistralDecoderLayer.forward / forward_xattn / flash_cross_attention_forward calls; and the
WHOLE prefill vs DattnMistralForCausalLM.forward -> prepare_inputs_labels_for_multimodal / encode_video_images /
encode_video_audios -> DattnMistralModel.forward run unmodified with use_cache=False (encoder outputs and logits) -- all to
<= 5e-5.  Not exercised: the use_cache=True branch.
Differences from Vidi1.5: no sqrt(D) normaliser, Mistral RMSNorm (w * x_hat, eps from config), one
MLP norm (= post_attention_layernorm), SwiGLU, no soft-caps, scale 1/sqrt(128), diagonal update without a norm,
residual added after summing the three attentions, learned-conv pooling to pool^2 tokens per frame, audio pool keeps
d_model, untied lm_head.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .vidi15_ref import (apply_rope, attend, linear, mm_norm, pos_embed, projector, rope_cos_sin, siglip_tower,
                         strip_image_token, whisper_encoder, xhat)


def mistral_norm(x, w, eps):
    return w.float() * xhat(x, eps)


def conv2d_pool_7b(x, w, s_out: int):
    """Conv2d(k=ceil(s_in/s_out), stride 1, no bias) then bilinear(align_corners=True) to s_out (pool.py:20-26)."""
    x = F.conv2d(x.float(), w.float())
    assert x.shape[-1] >= s_out
    return F.interpolate(x, size=s_out, mode="bilinear", align_corners=True)


def encode_video_images(sd, cfg, images):
    D, eps, s = cfg.llm.hidden, cfg.mm_eps, cfg.mm_image_pool_size
    feats = siglip_tower(sd, cfg, images)
    Fr, side = feats.shape[0], cfg.vis.side
    feats = feats.reshape(Fr, side, side, -1).permute(0, 3, 1, 2)
    x = conv2d_pool_7b(feats, sd["model.mm_rand_img_pool.conv.weight"], s).permute(0, 2, 3, 1)      # [F,s,s,dv]
    x = projector(sd, "model.mm_rand_img_projector", x)
    x = mm_norm(x, sd["model.mm_rand_img_norm.weight"], eps)
    x = x + xhat(pos_embed(sd, "model.mm_rand_pos_h", s, s, D), eps)[None, :, None, :]
    x = x + xhat(pos_embed(sd, "model.mm_rand_pos_w", s, s, D), eps)[None, None, :, :]
    x = x + xhat(pos_embed(sd, "model.mm_rand_pos_t", Fr, cfg.mm_time_interval, D), eps)[:, None, None, :]
    x = x.flatten(0, 2)
    mask = (x.abs().sum(-1) != 0) & bool(images.abs().sum() != 0)
    return mm_norm(x, sd["model.mm_rand_llm_norm.weight"], eps) * mask[:, None], mask


def encode_video_audios(sd, cfg, mels, audio_size):
    D, eps = cfg.llm.hidden, cfg.mm_eps
    a = whisper_encoder(sd, cfg, mels)
    s1 = int(math.floor(audio_size * (cfg.aud.max_source_positions / cfg.aud.nb_max_frames)))
    a = a.flatten(0, 1)[:s1]
    a = F.conv1d(a.t()[None], sd["model.mm_rand_aud_pool.weight"].float(), None, stride=cfg.mm_audio_pool_size)[0].t()
    s2 = int(math.floor(s1 / cfg.mm_audio_pool_size))
    a = projector(sd, "model.mm_rand_aud_projector", a[:s2])
    a = mm_norm(a, sd["model.mm_rand_aud_norm.weight"], eps)
    a = a + xhat(pos_embed(sd, "model.mm_rand_pos_t", s2, cfg.mm_time_interval, D), eps)
    mask = (a.abs().sum(-1) != 0) & bool(mels.abs().sum() != 0)
    return mm_norm(a, sd["model.mm_rand_llm_norm.weight"], eps) * mask[:, None], mask


def swiglu(x, sd, p):
    return linear(F.silu(linear(x, sd[f"{p}.mlp.gate_proj.weight"])) * linear(x, sd[f"{p}.mlp.up_proj.weight"]),
                  sd[f"{p}.mlp.down_proj.weight"])


def feed_forward(x, sd, p, eps):
    """mistral.py:131-137: x + mlp(post_attention_layernorm(x))."""
    return x + swiglu(mistral_norm(x, sd[f"{p}.post_attention_layernorm.weight"], eps), sd, p)


def stream_layer(S, sd, p, cfg):
    c = cfg.llm
    s = mistral_norm(S, sd[f"{p}.input_layernorm.weight"], c.rms_eps)
    K = linear(s, sd[f"{p}.self_attn.k_proj.weight"])
    V = linear(s, sd[f"{p}.self_attn.v_proj.weight"])
    Vrep = V.view(-1, c.kv_heads, c.head_dim).repeat_interleave(c.groups, 1).reshape(-1, c.q_dim)
    S = S + linear(Vrep, sd[f"{p}.self_attn.o_proj.weight"])             # no norm (mistral.py:223-225)
    return feed_forward(S, sd, p, c.rms_eps), K, V


def text_layer(H, sd, p, cfg, cos, sin, kv_streams):
    c = cfg.llm
    T = H.shape[0]
    scale = c.head_dim ** -0.5
    h = mistral_norm(H, sd[f"{p}.input_layernorm.weight"], c.rms_eps)
    q = linear(h, sd[f"{p}.self_attn.q_proj.weight"]).view(T, c.heads, c.head_dim).transpose(0, 1)
    k = linear(h, sd[f"{p}.self_attn.k_proj.weight"]).view(T, c.kv_heads, c.head_dim).transpose(0, 1)
    v = linear(h, sd[f"{p}.self_attn.v_proj.weight"]).view(T, c.kv_heads, c.head_dim).transpose(0, 1)
    i = torch.arange(T)
    bias = torch.zeros(T, T).masked_fill(~(i[None, :] <= i[:, None]), float("-inf"))
    parts = [attend(apply_rope(q, cos, sin), apply_rope(k, cos, sin), v, scale, None, bias)]
    for (K, V, mask) in kv_streams:
        Kh = K.view(-1, c.kv_heads, c.head_dim).transpose(0, 1)
        Vh = V.view(-1, c.kv_heads, c.head_dim).transpose(0, 1)
        any_valid = bool(mask.any())
        m = mask if any_valid else torch.ones_like(mask)
        kb = torch.zeros(m.shape[0]).masked_fill(~m, float("-inf"))[None, :]
        parts.append(attend(q, Kh, Vh, scale, None, kb) * float(any_valid))
    Wo = sd[f"{p}.self_attn.o_proj.weight"]
    H = H + sum(linear(x, Wo) for x in parts)                           # mistral.py:263
    return feed_forward(H, sd, p, c.rms_eps)


@torch.no_grad()
def prefill(sd, cfg, input_ids, images, mels, audio_size, return_intermediates=False, logits_to_keep: int = 0):
    c = cfg.llm
    ids = strip_image_token(input_ids)
    T = ids.shape[0]
    H = sd["model.embed_tokens.weight"].float()[ids]
    streams, inter = [], {"kv": []}
    if images is not None:
        X, mX = encode_video_images(sd, cfg, images); streams.append([X, mX]); inter["image_embeds"] = X
    if mels is not None:
        A, mA = encode_video_audios(sd, cfg, mels, audio_size); streams.append([A, mA]); inter["audio_embeds"] = A
    cos, sin = rope_cos_sin(T, c.head_dim, c.rope_theta)
    for l in range(c.layers):
        p = f"model.layers.{l}"
        kvs = []
        for st in streams:
            S_next, K, V = stream_layer(st[0], sd, p, cfg)
            kvs.append((K, V, st[1])); st[0] = S_next
        if return_intermediates:
            inter["kv"].append([(K, V) for (K, V, _) in kvs])
        H = text_layer(H, sd, p, cfg, cos, sin, kvs)
    Hn = mistral_norm(H, sd["model.norm.weight"], c.rms_eps)
    if logits_to_keep:
        Hn = Hn[-logits_to_keep:]
    logits = linear(Hn, sd["lm_head.weight"])
    return (logits, inter) if return_intermediates else logits


def greedy_generate(sd, cfg, input_ids, images, mels, audio_size, max_new_tokens=8, eos_id=2, return_margins=False):
    """Greedy decode by full re-prefill (DattnMistralForCausalLM.generate with do_sample=False, mistral.py:622-681).  Test
    infrastructure: quadratic, for mini dims only."""
    ids = input_ids.clone()
    out, margins = [], []
    for _ in range(max_new_tokens):
        logit = prefill(sd, cfg, ids, images, mels, audio_size, logits_to_keep=1)[-1]
        t2 = logit.topk(2).values
        nxt = int(logit.argmax(-1))
        out.append(nxt); margins.append(float(t2[0] - t2[1]))
        if nxt == eos_id:
            break
        ids = torch.cat([ids, torch.tensor([nxt])])
    return (out, margins) if return_margins else out

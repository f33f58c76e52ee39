"""ORACLE -- test infrastructure only.  fp32 CPU restatement of the Vidi1.5-9B prefill.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker / reported baseline.  The product path
(``vidi_b200``) never imports it and fails loudly when its CUDA library is missing.

PARITY PINNING: the reference (bytedance/vidi @ fc30c87) ships no tests and no golden vectors, has no
CPU path, and does not import as-is under this image's transformers 5.5.0 (pins 4.50.0; flash-attn /
deepspeed hard requirements) -- SURVEY.md section 8(c).  The oracle is nevertheless PINNED against
outputs of the reference itself run here: ``tests/golden/make_golden.py`` imports the reference's
unmodified modules from /root/reference through an import shim (``tests/golden/ref_shim.py``: absent
packages stubbed, ``flash_attn_func`` replaced by an eager fp32 restatement of its semantics) and
commits the outputs of
  * Conv2DPool / space_to_depth / resize_by_tokens / LearnablePosEmbd / RMSNorm / rms_norm / MLP,
  * DattnMMMixin.encode_video_images / encode_video_audios (on HF SiglipVisionModel / WhisperEncoder),
  * DattnGemma2DecoderLayer.forward + forward_xattn + flash_cross_attention_forward + splitted_call
    (two stacked layers, all three streams; the T2T half runs the installed HF Gemma2Attention),
  * the whole model: DattnGemma2ForCausalLM.forward -> prepare_inputs_labels_for_multimodal -> DattnGemma2Model.forward
    (gemma.py:267-424,484-601; multimodal.py:339-451) run unmodified with use_cache=False on instances assembled around
    those layers: sentinel stripping, embedding, the sqrt(D) normaliser, the layer loop, final norm, lm_head, soft-cap,
which ``tests/test_oracle_golden.py`` checks this restatement against to <= 5e-5 (logits included).
Not exercised by the fixtures: the use_cache=True branch (HF 5.x removed the 4.50 cache classes it constructs); decode
steps of ``greedy_generate`` therefore rest on prefill parity plus the cache-equivalence test in tests/.

Everything is fp32; functions take an HF-layout ``state_dict`` (keys of SURVEY.md section 8b).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# norms  (vidi/model/mm_layer/norm.py:9-25 ; HF modeling_gemma2.py Gemma2RMSNorm)
# ----------------------------------------------------------------------------------------------
def xhat(x: torch.Tensor, eps: float) -> torch.Tensor:
    x = x.float()
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)


def gemma_norm(x, w, eps=1e-6):
    """Gemma2RMSNorm: x_hat * (1 + w), product in fp32."""
    return xhat(x, eps) * (1.0 + w.float())


def mm_norm(x, w, eps=1e-5):
    """vidi RMSNorm: w * rms_norm(x)  (norm.py:17-22)."""
    return w.float() * xhat(x, eps)


def layer_norm(x, w, b, eps):
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps)


def linear(x, w, b=None):
    return F.linear(x, w.float(), None if b is None else b.float())


# ----------------------------------------------------------------------------------------------
# SigLIP tower  (vidi/model/mm_vision/siglip.py:29-34 -> HF modeling_siglip.py embeddings+encoder)
# ----------------------------------------------------------------------------------------------
def _mha(x, sd, pre, heads, q="q_proj", k="k_proj", v="v_proj", o="out_proj"):
    """Plain bidirectional MHA with biases where present; scale = head_dim**-0.5."""
    B, S, D = x.shape
    dh = D // heads
    def proj(name):
        return linear(x, sd[f"{pre}.{name}.weight"], sd.get(f"{pre}.{name}.bias"))
    qh = proj(q).view(B, S, heads, dh).transpose(1, 2)
    kh = proj(k).view(B, S, heads, dh).transpose(1, 2)
    vh = proj(v).view(B, S, heads, dh).transpose(1, 2)
    att = torch.softmax((qh @ kh.transpose(-1, -2)) * dh ** -0.5, dim=-1)
    out = (att @ vh).transpose(1, 2).reshape(B, S, D)
    return linear(out, sd[f"{pre}.{o}.weight"], sd.get(f"{pre}.{o}.bias"))


def siglip_tower(sd, cfg, images: torch.Tensor, prefix="model.mm_vis.vision_model") -> torch.Tensor:
    """images [F,3,H,W] -> hidden_states[select_layer] == output of layer run_layers  [F,P,d]."""
    v = cfg.vis
    x = F.conv2d(images.float(), sd[f"{prefix}.embeddings.patch_embedding.weight"].float(),
                 sd[f"{prefix}.embeddings.patch_embedding.bias"].float(), stride=v.patch)
    x = x.flatten(2).transpose(1, 2)                                     # [F,P,d]
    x = x + sd[f"{prefix}.embeddings.position_embedding.weight"].float()[None]
    for l in range(v.run_layers):
        p = f"{prefix}.encoder.layers.{l}"
        h = layer_norm(x, sd[f"{p}.layer_norm1.weight"], sd[f"{p}.layer_norm1.bias"], v.eps)
        x = x + _mha(h, sd, f"{p}.self_attn", v.heads)
        h = layer_norm(x, sd[f"{p}.layer_norm2.weight"], sd[f"{p}.layer_norm2.bias"], v.eps)
        h = F.gelu(linear(h, sd[f"{p}.mlp.fc1.weight"], sd[f"{p}.mlp.fc1.bias"]), approximate="tanh")
        x = x + linear(h, sd[f"{p}.mlp.fc2.weight"], sd[f"{p}.mlp.fc2.bias"])
    return x


# ----------------------------------------------------------------------------------------------
# Whisper encoder  (vidi/model/mm_audio/whisper.py:26-27 -> HF WhisperEncoder.forward)
# ----------------------------------------------------------------------------------------------
def whisper_encoder(sd, cfg, mels: torch.Tensor, prefix="model.mm_aud.encoder") -> torch.Tensor:
    """mels [C,128,3000] -> [C,1500,d]."""
    a = cfg.aud
    x = F.gelu(F.conv1d(mels.float(), sd[f"{prefix}.conv1.weight"].float(), sd[f"{prefix}.conv1.bias"].float(), padding=1))
    x = F.gelu(F.conv1d(x, sd[f"{prefix}.conv2.weight"].float(), sd[f"{prefix}.conv2.bias"].float(), stride=2, padding=1))
    x = x.permute(0, 2, 1) + sd[f"{prefix}.embed_positions.weight"].float()[None, : x.shape[-1]]
    for l in range(a.layers):
        p = f"{prefix}.layers.{l}"
        h = layer_norm(x, sd[f"{p}.self_attn_layer_norm.weight"], sd[f"{p}.self_attn_layer_norm.bias"], a.eps)
        x = x + _mha(h, sd, f"{p}.self_attn", a.heads)
        h = layer_norm(x, sd[f"{p}.final_layer_norm.weight"], sd[f"{p}.final_layer_norm.bias"], a.eps)
        h = F.gelu(linear(h, sd[f"{p}.fc1.weight"], sd[f"{p}.fc1.bias"]))
        x = x + linear(h, sd[f"{p}.fc2.weight"], sd[f"{p}.fc2.bias"])
    return layer_norm(x, sd[f"{prefix}.layer_norm.weight"], sd[f"{prefix}.layer_norm.bias"], a.eps)


# ----------------------------------------------------------------------------------------------
# pooling / positional  (mm_vision/pool.py:23-32, utils.py:134-171, mm_vision/pos.py:11-65)
# ----------------------------------------------------------------------------------------------
def space_to_depth(x: torch.Tensor, m: int) -> torch.Tensor:
    """[B,C,H,W] -> [B,C*m*m,H/m,W/m], channel index c*m*m + dy*m + dx  (utils.py:143-150)."""
    B, C, H, W = x.shape
    x = x.reshape(B, C, H // m, m, W // m, m).permute(0, 1, 3, 5, 2, 4)
    return x.reshape(B, C * m * m, H // m, W // m)


def conv2d_pool(x: torch.Tensor, hw, m: int) -> torch.Tensor:
    """pad(0,1,0,1) -> bilinear(align_corners=False) if hw[0] != 28 -> space_to_depth (pool.py:23-32)."""
    x = F.pad(x, (0, 1, 0, 1), value=0.0)
    if hw[0] != 28:
        x = F.interpolate(x, size=tuple(hw), mode="bilinear", align_corners=False)
    return space_to_depth(x, m)


def sinusoid(p: torch.Tensor, d: int) -> torch.Tensor:
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float) * -(math.log(10000.0) / d))
    pe = torch.zeros(len(p), d)
    pe[:, 0::2] = torch.sin(p.float()[:, None] * div)
    pe[:, 1::2] = torch.cos(p.float()[:, None] * div)
    return pe


def pos_embed(sd, prefix: str, l: int, N: int, d: int) -> torch.Tensor:
    """LearnablePosEmbd.forward in eval mode: p = i/(l-1)*(N-1); fp32 MLP (pos.py:41-58)."""
    assert l > 1
    p = torch.arange(l, dtype=torch.float) / (l - 1) * (N - 1)
    pe = sinusoid(p, d)
    h = F.gelu(linear(pe, sd[f"{prefix}.mlp.0.weight"], sd[f"{prefix}.mlp.0.bias"]))
    return linear(h, sd[f"{prefix}.mlp.2.weight"], sd[f"{prefix}.mlp.2.bias"])


def projector(sd, prefix: str, x):
    """mlp2x_gelu: Linear -> erf GELU -> Linear (mm_layer/mlp.py:16-22)."""
    h = F.gelu(linear(x, sd[f"{prefix}.model.0.weight"], sd[f"{prefix}.model.0.bias"]))
    return linear(h, sd[f"{prefix}.model.2.weight"], sd[f"{prefix}.model.2.bias"])


# ----------------------------------------------------------------------------------------------
# encode_video_images / encode_video_audios  (multimodal.py:156-252), batch of 1 video
# ----------------------------------------------------------------------------------------------
def encode_video_images(sd, cfg, images: torch.Tensor):
    """images [F,3,H,W] -> (features [N_v,D] *before* the sqrt(D) scale, mask [N_v] bool)."""
    D, eps = cfg.llm.hidden, cfg.mm_eps
    feats = siglip_tower(sd, cfg, images)                                 # [F,P,dv]
    Fr = feats.shape[0]
    s = cfg.vis.side
    feats = feats.reshape(Fr, s, s, -1).permute(0, 3, 1, 2)               # [F,dv,27,27]
    hw = cfg.image_hw(Fr)
    x = conv2d_pool(feats, hw, cfg.mm_image_pool_size).permute(0, 2, 3, 1)  # [F,h',w',4dv]
    x = projector(sd, "model.mm_rand_img_projector", x)
    x = mm_norm(x, sd["model.mm_rand_img_norm.weight"], eps)
    hp, wp = x.shape[1], x.shape[2]
    x = x + xhat(pos_embed(sd, "model.mm_rand_pos_h", hp, cfg.mm_image_pool_size, D), eps)[None, :, None, :]
    x = x + xhat(pos_embed(sd, "model.mm_rand_pos_w", wp, cfg.mm_image_pool_size, D), eps)[None, None, :, :]
    x = x + xhat(pos_embed(sd, "model.mm_rand_pos_t", Fr, cfg.mm_time_interval, D), eps)[:, None, None, :]
    x = x.flatten(0, 2)
    mask = (x.abs().sum(-1) != 0) & bool(images.abs().sum() != 0)
    x = mm_norm(x, sd["model.mm_rand_llm_norm.weight"], eps) * mask[:, None]
    return x, mask


def encode_video_audios(sd, cfg, mels: torch.Tensor, audio_size: int):
    """mels [C,128,3000] -> (features [N_a,D], mask [N_a])."""
    D, eps = cfg.llm.hidden, cfg.mm_eps
    a = whisper_encoder(sd, cfg, mels)                                    # [C,1500,da]
    s1 = int(math.floor(audio_size * (cfg.aud.max_source_positions / cfg.aud.nb_max_frames)))
    a = a.flatten(0, 1)[:s1]                                              # [s1,da]
    a = F.conv1d(a.t()[None], sd["model.mm_rand_aud_pool.weight"].float(), None,
                 stride=cfg.mm_audio_pool_size)[0].t()                    # [floor(s1/5),D]
    s2 = int(math.floor(s1 / cfg.mm_audio_pool_size))
    a = a[:s2]
    a = projector(sd, "model.mm_rand_aud_projector", a)
    a = mm_norm(a, sd["model.mm_rand_aud_norm.weight"], eps)
    a = a + xhat(pos_embed(sd, "model.mm_rand_pos_t", s2, cfg.mm_time_interval, D), eps)
    mask = (a.abs().sum(-1) != 0) & bool(mels.abs().sum() != 0)
    a = mm_norm(a, sd["model.mm_rand_llm_norm.weight"], eps) * mask[:, None]
    return a, mask


# ----------------------------------------------------------------------------------------------
# Gemma2 text pieces  (HF modeling_gemma2.py rotary / eager attention with softcap)
# ----------------------------------------------------------------------------------------------
def rope_cos_sin(T: int, dh: int, theta: float):
    inv = 1.0 / (theta ** (torch.arange(0, dh, 2, dtype=torch.float) / dh))
    fr = torch.arange(T, dtype=torch.float)[:, None] * inv[None]
    emb = torch.cat([fr, fr], -1)
    return emb.cos(), emb.sin()


def apply_rope(x, cos, sin):
    """x [H,T,dh]; rotate_half convention."""
    h = x.shape[-1] // 2
    rot = torch.cat([-x[..., h:], x[..., :h]], -1)
    return x * cos[None] + rot * sin[None]


def softcap(s, cap):
    return s if cap is None else cap * torch.tanh(s / cap)


def attend(q, k, v, scale, cap, bias=None):
    """q [Hq,Tq,dh], k/v [Hkv,N,dh] (GQA via repeat_interleave, == HF repeat_kv) -> [Tq,Hq*dh]."""
    g = q.shape[0] // k.shape[0]
    k = k.repeat_interleave(g, 0)
    v = v.repeat_interleave(g, 0)
    s = softcap((q @ k.transpose(-1, -2)) * scale, cap)
    if bias is not None:
        s = s + bias
    p = torch.softmax(s, -1)
    return (p @ v).transpose(0, 1).reshape(q.shape[1], -1)


def gemma_mlp(x, sd, p):
    g = F.gelu(linear(x, sd[f"{p}.mlp.gate_proj.weight"]), approximate="tanh")
    return linear(g * linear(x, sd[f"{p}.mlp.up_proj.weight"]), sd[f"{p}.mlp.down_proj.weight"])


def feed_forward(x, sd, p, eps):
    """DattnGemma2DecoderLayer.feed_foward (gemma.py:116-123)."""
    h = gemma_norm(x, sd[f"{p}.pre_feedforward_layernorm.weight"], eps)
    return x + gemma_norm(gemma_mlp(h, sd, p), sd[f"{p}.post_feedforward_layernorm.weight"], eps)


def stream_layer(S, sd, p, cfg):
    """One decoder layer applied to an image/audio stream S [N,D] (gemma.py:183-202).
    Returns (S_next, K [N,kv_dim], V [N,kv_dim]) -- K,V are what the reference caches (gemma.py:61-63)."""
    c = cfg.llm
    s = gemma_norm(S, sd[f"{p}.input_layernorm.weight"], c.rms_eps)
    K = linear(s, sd[f"{p}.self_attn.k_proj.weight"])
    V = linear(s, sd[f"{p}.self_attn.v_proj.weight"])
    Vrep = V.view(-1, c.kv_heads, c.head_dim).repeat_interleave(c.groups, 1).reshape(-1, c.q_dim)
    S = S + gemma_norm(linear(Vrep, sd[f"{p}.self_attn.o_proj.weight"]),
                       sd[f"{p}.post_attention_layernorm.weight"], c.rms_eps)
    S = feed_forward(S, sd, p, c.rms_eps)
    return S, K, V


def text_layer(H, sd, p, cfg, layer_idx, cos, sin, kv_streams, return_parts=False):
    """Text-stream half of the layer (gemma.py:160-175,185-192,236-238).
    kv_streams: list of (K [N,kv_dim], V [N,kv_dim], mask [N] bool) for image then audio."""
    c = cfg.llm
    T = H.shape[0]
    scale = c.query_pre_attn_scalar ** -0.5
    h = gemma_norm(H, sd[f"{p}.input_layernorm.weight"], c.rms_eps)
    q = linear(h, sd[f"{p}.self_attn.q_proj.weight"]).view(T, c.heads, c.head_dim).transpose(0, 1)
    k = linear(h, sd[f"{p}.self_attn.k_proj.weight"]).view(T, c.kv_heads, c.head_dim).transpose(0, 1)
    v = linear(h, sd[f"{p}.self_attn.v_proj.weight"]).view(T, c.kv_heads, c.head_dim).transpose(0, 1)
    # T2T: RoPE, causal, softcap; sliding window on even layers (gemma.py:104,153-158)
    i = torch.arange(T)
    allowed = i[None, :] <= i[:, None]
    if layer_idx % 2 == 0:
        allowed = allowed & (i[:, None] - i[None, :] < c.sliding_window)
    bias = torch.zeros(T, T).masked_fill(~allowed, float("-inf"))
    a = attend(apply_rope(q, cos, sin), apply_rope(k, cos, sin), v, scale, c.attn_softcap, bias)
    parts = [a]
    # T2V / T2A: no RoPE, non-causal, key-padding mask (gemma.py:58,81-91; xattn.py)
    for (K, V, mask) in kv_streams:
        Kh = K.view(-1, c.kv_heads, c.head_dim).transpose(0, 1)
        Vh = V.view(-1, c.kv_heads, c.head_dim).transpose(0, 1)
        any_valid = bool(mask.any())
        m = mask if any_valid else torch.ones_like(mask)                  # gemma.py:180-182
        kb = torch.zeros(m.shape[0]).masked_fill(~m, float("-inf"))[None, :]
        ax = attend(q, Kh, Vh, scale, c.attn_softcap, kb)
        parts.append(ax * float(any_valid))                               # gemma.py:192
    Wo = sd[f"{p}.self_attn.o_proj.weight"]
    att = sum(linear(x, Wo) for x in parts)                               # gemma.py:94,236
    H = H + gemma_norm(att, sd[f"{p}.post_attention_layernorm.weight"], c.rms_eps)
    H = feed_forward(H, sd, p, c.rms_eps)
    return (H, parts) if return_parts else H


# ----------------------------------------------------------------------------------------------
# full prefill
# ----------------------------------------------------------------------------------------------
def strip_image_token(input_ids: torch.Tensor, image_token_index: int = -200) -> torch.Tensor:
    """prepare_inputs_labels_for_multimodal drops the sentinel; nothing is spliced in (multimodal.py:377-397)."""
    ids = input_ids.reshape(-1)
    assert int((ids == image_token_index).sum()) <= 1, "only support at most one image for now."
    return ids[ids != image_token_index]


def normalizer(cfg, dtype=torch.float32) -> float:
    """torch.tensor(hidden**0.5, dtype=act) (gemma.py:353): the value is rounded to the activation dtype."""
    return float(torch.tensor(cfg.llm.hidden ** 0.5, dtype=dtype).float())


@torch.no_grad()
def prefill(sd, cfg, input_ids, images: Optional[torch.Tensor], mels: Optional[torch.Tensor],
            audio_size: Optional[int], normalizer_dtype=torch.float32, return_intermediates=False,
            logits_to_keep: int = 0):
    """DattnGemma2ForCausalLM.forward for one sample (gemma.py:484-601, 267-424).
    input_ids may contain the -200 sentinel.  Returns logits [T or k, vocab] (fp32) and, optionally,
    a dict of intermediates used by the per-stage GPU parity tests."""
    c = cfg.llm
    inter = {}
    ids = strip_image_token(input_ids)
    T = ids.shape[0]
    nrm = normalizer(cfg, normalizer_dtype)
    H = sd["model.embed_tokens.weight"].float()[ids] * nrm
    streams = []
    if images is not None:
        X, mX = encode_video_images(sd, cfg, images)
        inter["image_embeds"], inter["image_mask"] = X, mX
        streams.append([X * nrm, mX])
    if mels is not None:
        A, mA = encode_video_audios(sd, cfg, mels, audio_size)
        inter["audio_embeds"], inter["audio_mask"] = A, mA
        streams.append([A * nrm, mA])
    cos, sin = rope_cos_sin(T, c.head_dim, c.rope_theta)
    inter["kv"] = []
    for l in range(c.layers):
        p = f"model.layers.{l}"
        kvs = []
        for st in streams:
            S_next, K, V = stream_layer(st[0], sd, p, cfg)
            kvs.append((K, V, st[1]))
            st[0] = S_next
        if return_intermediates:
            inter["kv"].append([(K, V) for (K, V, _) in kvs])
        H = text_layer(H, sd, p, cfg, l, cos, sin, kvs)
        if return_intermediates:
            inter.setdefault("text_hidden", []).append(H)
    Hn = gemma_norm(H, sd["model.norm.weight"], c.rms_eps)
    if logits_to_keep:
        Hn = Hn[-logits_to_keep:]
    W = sd["model.embed_tokens.weight"] if c.tie_word_embeddings else sd["lm_head.weight"]
    logits = softcap(linear(Hn, W), c.final_softcap)
    if return_intermediates:
        inter["streams"] = [s[0] for s in streams]
        return logits, inter
    return logits


@torch.no_grad()
def greedy_generate(sd, cfg, input_ids, images, mels, audio_size, max_new_tokens=8, eos_id=107,
                    normalizer_dtype=torch.float32, return_margins=False):
    """Greedy decode by full re-prefill of the text (the streams do not depend on text, so the
    image/audio K,V are computed once).  Matches generate(do_sample=False) (gemma.py:603-655)."""
    c = cfg.llm
    ids = strip_image_token(input_ids).clone()
    nrm = normalizer(cfg, normalizer_dtype)
    streams = []
    if images is not None:
        X, mX = encode_video_images(sd, cfg, images); streams.append([X * nrm, mX])
    if mels is not None:
        A, mA = encode_video_audios(sd, cfg, mels, audio_size); streams.append([A * nrm, mA])
    kv_layers = []
    for l in range(c.layers):
        p = f"model.layers.{l}"
        kvs = []
        for st in streams:
            S_next, K, V = stream_layer(st[0], sd, p, cfg)
            kvs.append((K, V, st[1])); st[0] = S_next
        kv_layers.append(kvs)
    W = sd["model.embed_tokens.weight"] if c.tie_word_embeddings else sd["lm_head.weight"]
    out, margins = [], []       # margins: top-1 minus top-2 logit of every step (tests pick fixtures with decisive margins)
    for _ in range(max_new_tokens):
        T = ids.shape[0]
        H = sd["model.embed_tokens.weight"].float()[ids] * nrm
        cos, sin = rope_cos_sin(T, c.head_dim, c.rope_theta)
        for l in range(c.layers):
            H = text_layer(H, sd, f"model.layers.{l}", cfg, l, cos, sin, kv_layers[l])
        logit = softcap(linear(gemma_norm(H[-1:], sd["model.norm.weight"], c.rms_eps), W), c.final_softcap)
        nxt = int(logit.argmax(-1))
        out.append(nxt)
        t2 = logit[0].topk(2).values
        margins.append(float(t2[0] - t2[1]))
        if nxt == eos_id:
            break
        ids = torch.cat([ids, torch.tensor([nxt])])
    return (out, margins) if return_margins else out

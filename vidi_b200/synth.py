"""Seeded synthetic checkpoint in the reference's HF key layout (used by tests, smoke and bench).

No real checkpoint is reachable offline, so every test / bench uses a synthetic ``state_dict`` whose
keys and shapes are exactly those a Vidi1.5-9B checkpoint has (module tree of
Vidi1.5_9B/vidi/model/lmm/dattn/multimodal.py:44-94 + HF Gemma2 / SigLIP / Whisper).  Scales keep
activations sane through the full depth: linears N(0, 0.02^2), norm weights ~0 (Gemma ``1+w``) or ~1,
``mm_rand_llm_norm.weight`` ~ mm_std (finetune.sh:25).

Each tensor gets its own generator seeded from (seed, crc32(key)) so the dict is independent of
creation order and can be produced piecewise on any device.
"""
from __future__ import annotations

import zlib

import torch


def _gen(seed: int, key: str, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def tensor_specs(cfg) -> dict:
    """key -> (shape, kind, scale, dtype_is_fp32)."""
    L, V, A = cfg.llm, cfg.vis, cfg.aud
    D = L.hidden
    s = {}
    def lin(key, out, inp, bias=False, std=0.02, fp32=False):
        s[f"{key}.weight"] = ((out, inp), "normal", std, fp32)
        if bias:
            s[f"{key}.bias"] = ((out,), "normal", 0.02, fp32)
    def norm1(key, n, bias=False):            # weight ~ 1
        s[f"{key}.weight"] = ((n,), "one", 0.05, False)
        if bias:
            s[f"{key}.bias"] = ((n,), "normal", 0.02, False)
    def norm0(key, n):                        # gemma (1+w): weight ~ 0
        s[f"{key}.weight"] = ((n,), "normal", 0.05, False)

    s["model.embed_tokens.weight"] = ((L.vocab, D), "normal", 0.02, False)
    if not L.tie_word_embeddings:
        s["lm_head.weight"] = ((L.vocab, D), "normal", 0.02, False)
    for l in range(L.layers):
        p = f"model.layers.{l}"
        lin(f"{p}.self_attn.q_proj", L.q_dim, D)
        lin(f"{p}.self_attn.k_proj", L.kv_dim, D)
        lin(f"{p}.self_attn.v_proj", L.kv_dim, D)
        lin(f"{p}.self_attn.o_proj", D, L.q_dim)
        lin(f"{p}.mlp.gate_proj", L.inter, D)
        lin(f"{p}.mlp.up_proj", L.inter, D)
        lin(f"{p}.mlp.down_proj", D, L.inter)
        for n in ("input_layernorm", "post_attention_layernorm", "pre_feedforward_layernorm",
                  "post_feedforward_layernorm"):
            if hasattr(L, "final_softcap") or n in ("input_layernorm", "post_attention_layernorm"):
                (norm0 if hasattr(L, "final_softcap") else norm1)(f"{p}.{n}", D)
    (norm0 if hasattr(L, "final_softcap") else norm1)("model.norm", D)

    # SigLIP
    pv = "model.mm_vis.vision_model"
    s[f"{pv}.embeddings.patch_embedding.weight"] = ((V.hidden, 3, V.patch, V.patch), "normal", 0.02, False)
    s[f"{pv}.embeddings.patch_embedding.bias"] = ((V.hidden,), "normal", 0.02, False)
    s[f"{pv}.embeddings.position_embedding.weight"] = ((V.patches, V.hidden), "normal", 0.02, False)
    for l in range(V.layers):
        p = f"{pv}.encoder.layers.{l}"
        norm1(f"{p}.layer_norm1", V.hidden, bias=True)
        norm1(f"{p}.layer_norm2", V.hidden, bias=True)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lin(f"{p}.self_attn.{n}", V.hidden, V.hidden, bias=True, std=0.03)
        lin(f"{p}.mlp.fc1", V.inter, V.hidden, bias=True, std=0.03)
        lin(f"{p}.mlp.fc2", V.hidden, V.inter, bias=True, std=0.02)
    norm1(f"{pv}.post_layernorm", V.hidden, bias=True)           # present in ckpt, unused (select_layer=-2)

    # Whisper encoder
    pa = "model.mm_aud.encoder"
    s[f"{pa}.conv1.weight"] = ((A.d_model, A.mels, 3), "normal", 0.05, False)
    s[f"{pa}.conv1.bias"] = ((A.d_model,), "normal", 0.02, False)
    s[f"{pa}.conv2.weight"] = ((A.d_model, A.d_model, 3), "normal", 0.02, False)
    s[f"{pa}.conv2.bias"] = ((A.d_model,), "normal", 0.02, False)
    s[f"{pa}.embed_positions.weight"] = ((A.max_source_positions, A.d_model), "normal", 0.02, False)
    for l in range(A.layers):
        p = f"{pa}.layers.{l}"
        norm1(f"{p}.self_attn_layer_norm", A.d_model, bias=True)
        norm1(f"{p}.final_layer_norm", A.d_model, bias=True)
        lin(f"{p}.self_attn.q_proj", A.d_model, A.d_model, bias=True, std=0.03)
        lin(f"{p}.self_attn.k_proj", A.d_model, A.d_model, bias=False, std=0.03)
        lin(f"{p}.self_attn.v_proj", A.d_model, A.d_model, bias=True, std=0.03)
        lin(f"{p}.self_attn.out_proj", A.d_model, A.d_model, bias=True, std=0.03)
        lin(f"{p}.fc1", A.ffn, A.d_model, bias=True, std=0.03)
        lin(f"{p}.fc2", A.d_model, A.ffn, bias=True, std=0.02)
    norm1(f"{pa}.layer_norm", A.d_model, bias=True)

    # mm_rand_* glue
    s["model.mm_rand_llm_norm.weight"] = ((D,), "one", 0.05 , False)   # scaled by mm_std below
    if hasattr(cfg, "image_hw"):         # Vidi1.5
        pin = V.hidden * cfg.mm_image_pool_size ** 2
        s["model.mm_rand_aud_pool.weight"] = ((D, A.d_model, cfg.mm_audio_pool_size), "normal", 0.02, False)
        aud_proj_in = D
    else:                                # Vidi-7B: learned conv pool, audio pool keeps d_model
        import math
        k = math.ceil(V.side / cfg.mm_image_pool_size)
        s["model.mm_rand_img_pool.conv.weight"] = ((V.hidden, V.hidden, k, k), "normal", 0.02 / k, False)
        s["model.mm_rand_aud_pool.weight"] = ((A.d_model, A.d_model, cfg.mm_audio_pool_size), "normal", 0.02, False)
        pin = V.hidden
        aud_proj_in = A.d_model
    lin("model.mm_rand_img_projector.model.0", D, pin, bias=True)
    lin("model.mm_rand_img_projector.model.2", D, D, bias=True)
    norm1("model.mm_rand_img_norm", D)
    lin("model.mm_rand_aud_projector.model.0", D, aud_proj_in, bias=True)
    lin("model.mm_rand_aud_projector.model.2", D, D, bias=True)
    norm1("model.mm_rand_aud_norm", D)
    for n in ("h", "w", "t"):
        lin(f"model.mm_rand_pos_{n}.mlp.0", D, D, bias=True, std=0.03, fp32=True)
        lin(f"model.mm_rand_pos_{n}.mlp.2", D, D, bias=True, std=0.03, fp32=True)
    return s


def make_tensor(cfg, key, spec, seed=1234, device="cpu", dtype=torch.float32):
    shape, kind, scale, fp32 = spec
    g = _gen(seed, key, device)
    t = torch.randn(shape, generator=g, device=device, dtype=torch.float32) * scale
    if kind == "one":
        t = t + 1.0
    if key == "model.mm_rand_llm_norm.weight":
        t = t * cfg.mm_std
    return t if fp32 else t.to(dtype)


def make_state_dict(cfg, seed: int = 1234, device="cpu", dtype=torch.float32) -> dict:
    """Full HF-layout state_dict.  fp32 for the oracle; bf16 + device='cuda' for the engine/bench.
    (pos-MLP weights stay fp32 regardless, as in the reference: pos.py:38.)"""
    return {k: make_tensor(cfg, k, spec, seed, device, dtype) for k, spec in tensor_specs(cfg).items()}


def make_inputs(cfg, n_frames: int, n_chunks: int, n_text: int = 32, seed: int = 4321,
                audio_size: int | None = None, device="cpu"):
    """Synthetic inputs of BASELINE.md section 3: images randn.clamp(-1,1) [F,3,S,S], mels 0.5*randn
    [C,128,3000], ids = bos + random ids with the -200 sentinel at position 1 (as the chat template
    places ``<image>`` right after the turn header)."""
    g = torch.Generator(device=device); g.manual_seed(seed)
    images = torch.randn(n_frames, 3, cfg.vis.image, cfg.vis.image, generator=g, device=device).clamp_(-1, 1)
    mels = 0.5 * torch.randn(n_chunks, cfg.aud.mels, cfg.aud.nb_max_frames, generator=g, device=device)
    ids = torch.randint(3, cfg.llm.vocab, (n_text + 1,), generator=g, device=device)
    ids[0] = 2
    ids[1] = -200
    if audio_size is None:
        audio_size = min(n_frames * 100, n_chunks * cfg.aud.nb_max_frames)
    return ids, images, mels, int(audio_size)

"""HF-layout ``state_dict`` -> device-resident packed weights for the sm_100a kernels.

Key layout is the reference's.  Repacking done once at load:
  * q|k|v concatenated (text stream GEMM) -- the k|v rows double as the stream K||V projection (K12);
  * ``o_proj`` folded over the GQA repeat for the diagonal V2V update (K13):
        W_o' = W_o.view(D, Hkv, G, dh).sum(2)   (fp32 sum, rounded once)
    because ``repeat_kv`` is a repeat_interleave of KV heads (gemma.py:77-78, 195-197);
  * gate/up interleaved per 256-column GEMM tile so the GeGLU product is formed in the epilogue (K14);
  * conv weights flattened to GEMM form; projector columns permuted to the pooling kernel's (q, c) order;
  * positional-MLP weights split into 3 bf16 terms (see glue.cu) -- they stay effectively fp32 (pos.py:38);
  * biases / LayerNorm affine kept in fp32.
"""
from __future__ import annotations

import math
from types import SimpleNamespace as NS

import torch

BF16 = torch.bfloat16


def pack_glu(wg: torch.Tensor, wu: torch.Tensor, block_n: int = 256) -> torch.Tensor:
    """[I,K],[I,K] -> [2I,K] with rows grouped per GEMM tile as [block_n/2 gate | block_n/2 up]."""
    I, K = wg.shape
    h = block_n // 2
    assert I % h == 0, f"intermediate size {I} must be a multiple of {h}"
    return torch.stack([wg.view(I // h, h, K), wu.view(I // h, h, K)], dim=1).reshape(2 * I, K).contiguous()


def fold_o_proj(wo: torch.Tensor, kv_heads: int, groups: int, head_dim: int) -> torch.Tensor:
    D = wo.shape[0]
    return wo.float().view(D, kv_heads, groups, head_dim).sum(2).reshape(D, kv_heads * head_dim)


def _dev(t, device, dtype=None):
    return t.to(device=device, dtype=dtype if dtype is not None else t.dtype, non_blocking=True).contiguous()


def load_llm_layer(sd, l: int, c, device, glu_block: int = 256, pop: bool = False):
    g = sd.pop if pop else sd.__getitem__
    p = f"model.layers.{l}"
    wq, wk, wv = g(f"{p}.self_attn.q_proj.weight"), g(f"{p}.self_attn.k_proj.weight"), g(f"{p}.self_attn.v_proj.weight")
    wo = g(f"{p}.self_attn.o_proj.weight")
    L = NS()
    L.wqkv = _dev(torch.cat([wq, wk, wv], 0), device, BF16)
    L.wkv = L.wqkv[c.q_dim:]                                        # rows k|v (contiguous view)
    L.wo = _dev(wo, device, BF16)
    L.wo_fold = _dev(fold_o_proj(wo.to(device), c.kv_heads, c.groups, c.head_dim), device, BF16)
    L.wgu = pack_glu(_dev(g(f"{p}.mlp.gate_proj.weight"), device, BF16), _dev(g(f"{p}.mlp.up_proj.weight"), device, BF16),
                     glu_block)
    L.wd = _dev(g(f"{p}.mlp.down_proj.weight"), device, BF16)
    L.n_in = _dev(g(f"{p}.input_layernorm.weight"), device, BF16)
    L.n_post = _dev(g(f"{p}.post_attention_layernorm.weight"), device, BF16)
    if f"{p}.pre_feedforward_layernorm.weight" in sd:
        L.n_preff = _dev(g(f"{p}.pre_feedforward_layernorm.weight"), device, BF16)
        L.n_postff = _dev(g(f"{p}.post_feedforward_layernorm.weight"), device, BF16)
    return L


def fold_layernorm(w: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """Linear(LayerNorm(x)) with the affine part of the LayerNorm folded into the linear layer:
         y = W (gamma * xhat + beta) + b = (W diag(gamma)) xhat + (W beta + b),      xhat = (x - mean) * rstd
       and, with W' = bf16(W diag(gamma)) applied to the RAW x by the GEMM,  y = rstd * (x W'^T - mean * colsum(W')) + b'.
    Returns (W' bf16 [N,K], colsum fp32 [N] of the ROUNDED W', b' fp32 [N])."""
    wf = w.float()
    wp = (wf * gamma.float()[None, :]).to(BF16)
    return wp, wp.float().sum(1), (b.float() + wf @ beta.float())


def load_tower_layer(sd, p: str, device, names, pop: bool = False):
    """names: dict(ln1, ln2, attn, fc1, fc2, out) -> key stems (SigLIP and Whisper differ only in names)."""
    g = sd.pop if pop else sd.__getitem__
    L = NS()
    a = names["attn"]
    wq, wk, wv = g(f"{p}.{a}.q_proj.weight"), g(f"{p}.{a}.k_proj.weight"), g(f"{p}.{a}.v_proj.weight")
    d = wq.shape[0]
    def bias(n):
        k = f"{p}.{a}.{n}.bias"
        return (g(k) if k in sd else torch.zeros(d, device=wq.device)).float()
    L.wqkv = _dev(torch.cat([wq, wk, wv], 0), device, BF16)
    L.bqkv = _dev(torch.cat([bias("q_proj"), bias("k_proj"), bias("v_proj")], 0), device, torch.float32)
    L.wo = _dev(g(f"{p}.{a}.{names['out']}.weight"), device, BF16)
    L.bo = _dev(g(f"{p}.{a}.{names['out']}.bias"), device, torch.float32)
    L.w1 = _dev(g(f"{p}.{names['fc1']}.weight"), device, BF16)
    L.b1 = _dev(g(f"{p}.{names['fc1']}.bias"), device, torch.float32)
    L.w2 = _dev(g(f"{p}.{names['fc2']}.weight"), device, BF16)
    L.b2 = _dev(g(f"{p}.{names['fc2']}.bias"), device, torch.float32)
    for tag, n in (("ln1", names["ln1"]), ("ln2", names["ln2"])):
        setattr(L, f"{tag}_w", _dev(g(f"{p}.{n}.weight"), device, torch.float32))
        setattr(L, f"{tag}_b", _dev(g(f"{p}.{n}.bias"), device, torch.float32))
    return L


def fold_tower_layer(L):
    """adds the LayerNorm-folded copies of the q|k|v and fc1 weights used by the optional ops.gemm_ln path (engine.fold_ln)"""
    L.wqkv_f, L.cqkv, L.bqkv_f = fold_layernorm(L.wqkv, L.bqkv, L.ln1_w, L.ln1_b)
    L.w1_f, L.c1, L.b1_f = fold_layernorm(L.w1, L.b1, L.ln2_w, L.ln2_b)
    return L


SIGLIP_NAMES = dict(ln1="layer_norm1", ln2="layer_norm2", attn="self_attn", out="out_proj", fc1="mlp.fc1", fc2="mlp.fc2")
WHISPER_NAMES = dict(ln1="self_attn_layer_norm", ln2="final_layer_norm", attn="self_attn", out="out_proj", fc1="fc1", fc2="fc2")


def _pos_mlp(sd, prefix, device, ops):
    P = NS()
    for i, tag in ((0, "0"), (2, "2")):
        w = _dev(sd[f"{prefix}.mlp.{i}.weight"], device, torch.float32)
        setattr(P, f"w{tag}", ops.split3(w, 1))                       # [D, 3D] bf16 = [hi | lo | hi]
        setattr(P, f"b{tag}", _dev(sd[f"{prefix}.mlp.{i}.bias"], device, torch.float32))
    return P


def load_vidi15(sd: dict, cfg, device, ops, pop: bool = False):
    """Pack every weight the Vidi1.5 prefill needs.  ``ops`` is vidi_b200.ops (split3 runs on the GPU).
    With pop=True entries are removed from ``sd`` as they are consumed (bounds peak memory at 9B scale)."""
    c, v, a = cfg.llm, cfg.vis, cfg.aud
    D = c.hidden
    W = NS()
    W.embed = _dev(sd["model.embed_tokens.weight"], device, BF16)
    W.lm_head = W.embed if c.tie_word_embeddings or "lm_head.weight" not in sd else _dev(sd["lm_head.weight"], device, BF16)
    W.final_norm = _dev(sd["model.norm.weight"], device, BF16)
    W.layers = [load_llm_layer(sd, l, c, device, pop=pop) for l in range(c.layers)]

    # SigLIP
    pv = "model.mm_vis.vision_model"
    kp = 3 * v.patch * v.patch
    kpad = ((kp + 63) // 64) * 64
    wpe = torch.zeros(v.hidden, kpad, dtype=torch.float32, device=device)
    wpe[:, :kp] = sd[f"{pv}.embeddings.patch_embedding.weight"].to(device).float().reshape(v.hidden, kp)
    W.vis = NS(patch_w=wpe.to(BF16), kpad=kpad,
               patch_b=_dev(sd[f"{pv}.embeddings.patch_embedding.bias"], device, torch.float32),
               pos=_dev(sd[f"{pv}.embeddings.position_embedding.weight"], device, BF16),
               layers=[load_tower_layer(sd, f"{pv}.encoder.layers.{l}", device, SIGLIP_NAMES, pop=pop)
                       for l in range(v.run_layers)])
    # Whisper encoder
    pa = "model.mm_aud.encoder"
    W.aud = NS(conv1_w=_dev(sd[f"{pa}.conv1.weight"].permute(0, 2, 1).reshape(a.d_model, 3 * a.mels), device, BF16),
               conv1_b=_dev(sd[f"{pa}.conv1.bias"], device, torch.float32),
               conv2_w=_dev(sd[f"{pa}.conv2.weight"].permute(0, 2, 1).reshape(a.d_model, 3 * a.d_model), device, BF16),
               conv2_b=_dev(sd[f"{pa}.conv2.bias"], device, torch.float32),
               pos=_dev(sd[f"{pa}.embed_positions.weight"], device, BF16),
               ln_w=_dev(sd[f"{pa}.layer_norm.weight"], device, torch.float32),
               ln_b=_dev(sd[f"{pa}.layer_norm.bias"], device, torch.float32),
               layers=[load_tower_layer(sd, f"{pa}.layers.{l}", device, WHISPER_NAMES, pop=pop) for l in range(a.layers)])
    # mm glue
    m = cfg.mm_image_pool_size
    w1 = sd["model.mm_rand_img_projector.model.0.weight"]
    if "model.mm_rand_img_pool.conv.weight" in sd:                              # Vidi-7B learned pool: conv as GEMM
        wc = sd["model.mm_rand_img_pool.conv.weight"]                           # [dv, dv, k, k] -> columns (ky*k+kx)*dv + c
        W.img_pool_k = wc.shape[-1]
        W.img_pool_w = _dev(wc.permute(0, 2, 3, 1).reshape(wc.shape[0], -1), device, BF16)
    else:                                                                       # Vidi1.5: [D, dv*m*m], columns c*m*m + q
        w1 = w1.reshape(D, v.hidden, m * m).permute(0, 2, 1).reshape(D, m * m * v.hidden)   # -> q*dv + c
    W.img_proj = NS(w1=_dev(w1, device, BF16), b1=_dev(sd["model.mm_rand_img_projector.model.0.bias"], device, torch.float32),
                    w2=_dev(sd["model.mm_rand_img_projector.model.2.weight"], device, BF16),
                    b2=_dev(sd["model.mm_rand_img_projector.model.2.bias"], device, torch.float32))
    wp = sd["model.mm_rand_aud_pool.weight"]                                    # [D, da, k]
    W.aud_pool = _dev(wp.permute(0, 2, 1).reshape(wp.shape[0], -1), device, BF16)       # columns j*da + c
    W.aud_proj = NS(w1=_dev(sd["model.mm_rand_aud_projector.model.0.weight"], device, BF16),
                    b1=_dev(sd["model.mm_rand_aud_projector.model.0.bias"], device, torch.float32),
                    w2=_dev(sd["model.mm_rand_aud_projector.model.2.weight"], device, BF16),
                    b2=_dev(sd["model.mm_rand_aud_projector.model.2.bias"], device, torch.float32))
    W.img_norm = _dev(sd["model.mm_rand_img_norm.weight"], device, BF16)
    W.aud_norm = _dev(sd["model.mm_rand_aud_norm.weight"], device, BF16)
    W.llm_norm = _dev(sd["model.mm_rand_llm_norm.weight"], device, BF16)
    W.pos = {n: _pos_mlp(sd, f"model.mm_rand_pos_{n}", device, ops) for n in ("h", "w", "t")}
    # host-computed constants, exactly as the reference builds them on the CPU at init (pos.py:14-16; HF rotary)
    W.div_term = torch.exp(torch.arange(0, D, 2, dtype=torch.float) * -(math.log(10000.0) / D)).to(device)
    W.inv_freq = (1.0 / (c.rope_theta ** (torch.arange(0, c.head_dim, 2, dtype=torch.float) / c.head_dim))).to(device)
    return W

"""Tensor-level wrappers over the C ABI.  torch is used for device memory and the current stream only;
every computation is a kernel from ``libvidi_b200.so``.  All wrappers raise on CPU tensors."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import lib as _lib

ACT_NONE, ACT_GELU_ERF, ACT_GELU_TANH, ACT_SOFTCAP, ACT_SILU = 0, 1, 2, 3, 4
GLU_NONE, GLU_GELU_TANH, GLU_SILU = 0, 1, 2
BF16 = torch.bfloat16

# CTA-pair (cta_group::2) GEMM for large-M problems; flipped on once it beats the 1-CTA kernel on the box
USE_2CTA = True

# tower attention: exponentiate every ATTN_POLY-th score pair with the FMA-pipe polynomial instead of MUFU.EX2 (0 = MUFU only);
# validated on B200 in round 2 (tests/test_preprocess_gpu.py::test_attn_dense_poly_exp2_variants), default chosen by measurement
ATTN_POLY = 0

# bench instrumentation: when PROFILE is a list, every gemm() appends (tag, algorithmic_flops, start_evt, end_evt)
PROFILE = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("vidi_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def _rowmajor(t: torch.Tensor) -> int:
    """leading dimension (elements) of a 2-D row-major view; last dim must be contiguous."""
    assert t.dim() == 2 and t.stride(1) == 1, f"need 2-D row-major, got {tuple(t.shape)} {t.stride()}"
    return t.stride(0)


def pick_block_n(M: int, N: int, glu: bool = False) -> int:
    if glu:
        return 256
    if M <= 256 and N <= 16384:
        return 64                       # text stream (M ~ 32): weight-bandwidth bound, spread the N tiles over more SMs
    if N % 256 != 0 and N % 192 == 0 and N < 2048 and not (USE_2CTA and M >= 1024):
        return 192                      # 1-CTA kernel, N = 1152: exact tiling instead of a half-empty last 256 tile.  On the CTA-pair
                                        # kernel (M >= 1024) the 256-wide tile wins despite the padding: sustained tower block, N = 1152
                                        # sites 832 -> 874 / 1027 -> 1071 TF/s (profiles/r02_tower_ab.txt)
    if N >= 1024 or N % 256 == 0:
        return 256
    return 128 if N > 64 else 64


def ln_block_n(N: int) -> int:
    """column tile of the CTA-pair GEMM on the LayerNorm-folded tower sites (fixes the layout of the row-statistics buffer)"""
    return pick_block_n(1 << 20, N)


def ln_stats_parts(N: int) -> int:
    """(sum, sumsq) pairs per row written by gemm_ln(..., stats=) for an N-column output"""
    bn = ln_block_n(N)
    return 2 * ((N + bn - 1) // bn)


def gemm_ln(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
            res_mod: int = 0, act: int = ACT_NONE, out: Optional[torch.Tensor] = None, ln=None, stats: Optional[torch.Tensor] = None,
            tag: str = "gemm", alg_k: Optional[int] = None) -> torch.Tensor:
    """CTA-pair GEMM of a pre-LN tower block.  ln = (stats_in [M, parts, 2] fp32, colsum [N] fp32, eps): ``a`` is the raw
    residual stream and ``w`` / ``bias`` are the gamma / beta-folded weights (weights.fold_layernorm); stats: fp32
    [M, ln_stats_parts(N), 2] receiving the row statistics of the stored output (the next block's ``ln`` input)."""
    L = _lib.load()
    assert a.dtype == BF16 and w.dtype == BF16
    M, K = a.shape
    N, Kw = w.shape
    assert K == Kw, (a.shape, w.shape)
    block_n = ln_block_n(N)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=BF16)
    assert out.shape == (M, N) and out.dtype == BF16
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if residual is not None:
        assert residual.dtype == BF16
    ln_stats, ln_parts, ln_colsum, ln_eps = None, 0, None, 0.0
    if ln is not None:
        ln_stats, ln_colsum, ln_eps = ln
        assert ln_stats.dtype == torch.float32 and ln_stats.dim() == 3 and ln_stats.shape[0] == M and ln_stats.shape[2] == 2
        assert ln_stats.is_contiguous() and ln_colsum.dtype == torch.float32 and ln_colsum.numel() == N
        ln_parts = ln_stats.shape[1]
    if stats is not None:
        assert stats.dtype == torch.float32 and stats.shape == (M, ln_stats_parts(N), 2) and stats.is_contiguous()
    if M == 0:
        return out
    prof = PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.vidi_gemm_bf16_2cta_ln(_ptr(a), _rowmajor(a), _ptr(w), _rowmajor(w), _ptr(out), _rowmajor(out), M, N, K,
                                  _ptr(bias), _ptr(residual), _rowmajor(residual) if residual is not None else 0, res_mod,
                                  act, 0.0, block_n, _ptr(ln_stats), ln_parts, _ptr(ln_colsum), float(ln_eps), _ptr(stats),
                                  _stream())
    _lib.check(rc, "gemm_bf16_2cta_ln")
    if prof is not None:
        e1.record()
        prof.append((tag, 2.0 * M * N * (alg_k if alg_k is not None else K), e0, e1))
    return out


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         res_mod: int = 0, act: int = ACT_NONE, act_param: float = 0.0, out: Optional[torch.Tensor] = None,
         out_fp32: bool = False, glu: int = GLU_NONE, block_n: Optional[int] = None, tag: str = "gemm",
         alg_k: Optional[int] = None, cta2: Optional[bool] = None) -> torch.Tensor:
    """out[M,N(/2)] = epi(a[M,K] @ w[N,K]^T).  a,w bf16; bias fp32.
    tag / alg_k only feed the bench's roofline accounting (alg_k = un-padded contraction length)."""
    L = _lib.load()
    assert a.dtype == BF16 and w.dtype == BF16
    M, K = a.shape
    N, Kw = w.shape
    assert K == Kw, (a.shape, w.shape)
    n_out = N // 2 if glu else N
    if block_n is None:
        block_n = pick_block_n(M, N, bool(glu))
    if out is None:
        out = torch.empty(M, n_out, device=a.device, dtype=torch.float32 if out_fp32 else BF16)
    assert out.shape == (M, n_out)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if residual is not None:
        assert residual.dtype == BF16
    if M == 0:
        return out
    prof = PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    if cta2 is None:
        cta2 = USE_2CTA and M >= 1024 and block_n in (128, 192, 256)
    fn = L.vidi_gemm_bf16_2cta if cta2 else L.vidi_gemm_bf16
    rc = fn(_ptr(a), _rowmajor(a), _ptr(w), _rowmajor(w), _ptr(out), _rowmajor(out), M, N, K,
                          _ptr(bias), _ptr(residual), _rowmajor(residual) if residual is not None else 0, res_mod,
                          act, act_param, 1 if out.dtype == torch.float32 else 0, glu, block_n, _stream())
    _lib.check(rc, "gemm_bf16")
    if prof is not None:
        e1.record()
        prof.append((tag, 2.0 * M * N * (alg_k if alg_k is not None else K), e0, e1))
    return out


def rmsnorm(x, w, eps, add_one: bool, out=None, out_scale: float = 1.0):
    L = _lib.load()
    rows, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    _lib.check(L.vidi_rmsnorm(_ptr(x), _rowmajor(x), _ptr(w), _ptr(out), _rowmajor(out), rows, D, eps, int(add_one),
                              out_scale, _stream()), "rmsnorm")
    return out


def residual_norm(x, y, w_post, w_next, h, eps, post_mode: int, next_add_one: bool):
    """x += post(y); h = norm(x) (h may be None). In place on x."""
    L = _lib.load()
    rows, D = x.shape
    _lib.check(L.vidi_residual_norm(_ptr(x), _rowmajor(x), _ptr(y), _rowmajor(y), _ptr(w_post), _ptr(w_next),
                                    _ptr(h), _rowmajor(h) if h is not None else 0, rows, D, eps, post_mode,
                                    int(next_add_one), _stream()), "residual_norm")
    return x, h


def layernorm(x, w, b, eps, out=None):
    L = _lib.load()
    rows, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    assert w.dtype == torch.float32 and b.dtype == torch.float32
    _lib.check(L.vidi_layernorm(_ptr(x), _rowmajor(x), _ptr(w), _ptr(b), _ptr(out), _rowmajor(out), rows, D, eps,
                                _stream()), "layernorm")
    return out


def mm_finish(proj, w_mod, w_llm, tabs: Sequence[torch.Tensor], divs, mods, offs, n_offset: int, sample_valid: bool,
              normalizer: float, eps: float, out=None, mask=None):
    L = _lib.load()
    rows, D = proj.shape
    if out is None:
        out = torch.empty_like(proj)
    if mask is None:
        mask = torch.empty(rows, device=proj.device, dtype=torch.uint8)
    n = len(tabs)
    for t in tabs:
        assert t.dtype == torch.float32 and t.is_contiguous() and t.shape[-1] == D
    tp = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in tabs])
    ia = lambda v: (C.c_int * max(n, 1))(*[int(x) for x in v])
    _lib.check(L.vidi_mm_finish(_ptr(proj), _rowmajor(proj), _ptr(w_mod), _ptr(w_llm), tp, ia(divs), ia(mods), ia(offs),
                                n, n_offset, int(sample_valid), normalizer, _ptr(out), _rowmajor(out), _ptr(mask),
                                rows, D, eps, _stream()), "mm_finish")
    return out, mask


def rmsnorm_f32(x, eps, round_bf16: bool):
    L = _lib.load()
    rows, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty_like(x)
    _lib.check(L.vidi_rmsnorm_f32(_ptr(x), _ptr(y), rows, D, eps, int(round_bf16), _stream()), "rmsnorm_f32")
    return y


def patch_im2col(images, patch: int, kpad: int):
    L = _lib.load()
    F_, Cc, S, S2 = images.shape
    assert Cc == 3 and S == S2 and images.dtype == BF16 and images.is_contiguous()
    side = S // patch
    out = torch.empty(F_ * side * side, kpad, device=images.device, dtype=BF16)
    _lib.check(L.vidi_patch_im2col(_ptr(images), _ptr(out), F_, S, patch, kpad, _stream()), "patch_im2col")
    return out


def whisper_im2col1(mel):
    L = _lib.load()
    Cn, mels, T = mel.shape
    assert mel.dtype == BF16 and mel.is_contiguous()
    out = torch.empty(Cn * T, 3 * mels, device=mel.device, dtype=BF16)
    _lib.check(L.vidi_whisper_im2col1(_ptr(mel), _ptr(out), Cn, mels, T, _stream()), "whisper_im2col1")
    return out


def whisper_im2col2(x, Cn: int, T: int):
    L = _lib.load()
    d = x.shape[-1]
    assert x.dtype == BF16 and x.is_contiguous() and x.numel() == Cn * T * d
    out = torch.empty(Cn * (T // 2), 3 * d, device=x.device, dtype=BF16)
    _lib.check(L.vidi_whisper_im2col2(_ptr(x), _ptr(out), Cn, T, d, _stream()), "whisper_im2col2")
    return out


def pool_s2d(P, F_: int, side: int, h: int, w: int, m: int):
    L = _lib.load()
    d = P.shape[-1]
    assert P.dtype == BF16 and P.is_contiguous() and P.numel() == F_ * side * side * d
    out = torch.empty(F_ * (h // m) * (w // m), m * m * d, device=P.device, dtype=BF16)
    _lib.check(L.vidi_pool_s2d(_ptr(P), _ptr(out), F_, side, d, h, w, m, _stream()), "pool_s2d")
    return out


def conv_window_gather(P, F_: int, side: int, k: int):
    L = _lib.load()
    d = P.shape[-1]
    so = side - k + 1
    assert P.dtype == BF16 and P.is_contiguous() and P.numel() == F_ * side * side * d
    out = torch.empty(F_ * so * so, k * k * d, device=P.device, dtype=BF16)
    _lib.check(L.vidi_conv_window_gather(_ptr(P), _ptr(out), F_, side, d, k, _stream()), "conv_window_gather")
    return out


def bilinear_ac(X, F_: int, si: int, so: int):
    L = _lib.load()
    d = X.shape[-1]
    assert X.dtype == BF16 and X.is_contiguous() and X.numel() == F_ * si * si * d
    out = torch.empty(F_ * so * so, d, device=X.device, dtype=BF16)
    _lib.check(L.vidi_bilinear_ac(_ptr(X), _ptr(out), F_, si, so, d, _stream()), "bilinear_ac")
    return out


def embed_gather(ids, E, normalizer: float):
    L = _lib.load()
    assert ids.dtype == torch.int64 and ids.is_contiguous() and E.dtype == BF16 and E.is_contiguous()
    T = ids.numel()
    out = torch.empty(T, E.shape[1], device=E.device, dtype=BF16)
    _lib.check(L.vidi_embed_gather(_ptr(ids), _ptr(E), _ptr(out), T, E.shape[1], E.shape[0], normalizer, _stream()),
               "embed_gather")
    return out


def sinusoid_split(div_term, rows: int, i0: int, l: int, N: int, D: int):
    L = _lib.load()
    assert div_term.dtype == torch.float32 and div_term.numel() == D // 2
    out = torch.empty(rows, 3 * D, device=div_term.device, dtype=BF16)
    _lib.check(L.vidi_sinusoid_split(_ptr(div_term), _ptr(out), rows, i0, l, N, D, _stream()), "sinusoid_split")
    return out


def split3(x, mode: int):
    L = _lib.load()
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    rows, D = x.shape
    out = torch.empty(rows, 3 * D, device=x.device, dtype=BF16)
    _lib.check(L.vidi_split3(_ptr(x), _ptr(out), rows, D, mode, _stream()), "split3")
    return out


def cast_bf16(x):
    L = _lib.load()
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    _lib.check(L.vidi_cast_f32_bf16(_ptr(x), _ptr(y), x.numel(), _stream()), "cast_f32_bf16")
    return y


def attn_dense(qkv, B: int, S: int, H: int, dh: int, scale: float, out=None, impl: str = "auto"):
    """qkv [B*S, 3*H*dh] with Q|K|V sections -> out [B*S, H*dh].  impl: "auto" (tcgen05 for dh 64/72) | "v1" | "v2" (earlier tcgen05 generations) | "mma"."""
    L = _lib.load()
    d = H * dh
    assert qkv.dtype == BF16 and qkv.shape == (B * S, 3 * d)
    if out is None:
        out = torch.empty(B * S, d, device=qkv.device, dtype=BF16)
    if impl == "auto" and ATTN_POLY and dh in (64, 72) and S > 128:
        impl = f"poly{ATTN_POLY}"
    if impl.startswith("poly"):          # A/B only: FMA-pipe exp2 for every (2|3|4)-th score pair, e.g. impl="poly4"
        _lib.check(L.vidi_attn_dense_poly(_ptr(qkv), _rowmajor(qkv), _ptr(out), _rowmajor(out), B, S, H, dh, scale, int(impl[4:]),
                                          _stream()), "attn_dense_poly")
        return out
    fn = L.vidi_attn_dense_mma if impl == "mma" else L.vidi_attn_dense_v1 if impl == "v1" else L.vidi_attn_dense_v2 if impl == "v2" else L.vidi_attn_dense
    _lib.check(fn(_ptr(qkv), _rowmajor(qkv), 0, d, 2 * d, _ptr(out), _rowmajor(out), B, S, H, dh, scale, _stream()),
               "attn_dense")
    return out


def xattn_splits(n_keys: int, hkv: int, n_sms: int = 148) -> int:
    """number of key splits so that splits*hkv CTAs fill the SMs once (one persistent-style wave), >= 512 keys per split.
    Fewer, longer splits amortise the per-CTA prologue and keep the (O, LSE) partial set small for the merge."""
    if n_keys <= 0:
        return 1
    want = max(1, n_sms // max(hkv, 1))
    return max(1, min(want, (n_keys + 511) // 512))


def xattn_split_plan(n_keys: Sequence[int], hkv: int, n_sms: int = 148) -> list:
    """key splits for the segments of ONE launch: the total fills the SMs once (n_sms // hkv CTAs per KV head), shared between the
    segments in proportion to their key counts, at least 1 each and at most one split per 512 keys."""
    want = max(1, n_sms // max(hkv, 1))
    caps = [max(1, (max(n, 0) + 511) // 512) for n in n_keys]
    tot = sum(max(n, 0) for n in n_keys)
    if len(n_keys) == 1 or tot == 0:
        return [min(want, c) for c in caps]
    out = [max(1, min(c, int(round(want * max(n, 0) / tot)))) for n, c in zip(n_keys, caps)]
    while sum(out) > want and max(out) > 1:
        i = max(range(len(out)), key=lambda j: out[j])
        out[i] -= 1
    return out


def xattn_splitkv_seg(q, k, v, segs, splits, Hq: int, Hkv: int, dh: int, scale: float, softcap: float, opart, lse):
    """Both key segments of a layer in one call (one launch on the tcgen05 path).  k, v: [N, *] views of the layer's K||V cache from
    row 0; segs: [(row0, rows, kmask or None)]; opart fp32 flat [sum(splits) * T * Hq * dh], lse fp32 flat [sum(splits) * T * Hq]."""
    L = _lib.load()
    T, N = q.shape[0], k.shape[0]
    n = len(segs)
    assert 1 <= n <= 2 and len(splits) == n and q.dtype == BF16 and k.dtype == BF16 and v.dtype == BF16
    assert opart.dtype == torch.float32 and opart.numel() >= sum(splits) * T * Hq * dh and lse.numel() >= sum(splits) * T * Hq
    ia = lambda vals: (C.c_int32 * 2)(*(list(vals) + [0] * (2 - n)))
    masks = (C.c_void_p * 2)(*([(_ptr(m) if (m is not None and m.numel()) else None) for _, _, m in segs] + [None] * (2 - n)))
    _lib.check(L.vidi_xattn_splitkv_seg(_ptr(q), _rowmajor(q), _ptr(k) if N else None, _ptr(v) if N else None, k.stride(0) if N else 8, N, n,
                                        ia([s[0] for s in segs]), ia([s[1] for s in segs]), ia(splits), masks, T, Hq, Hkv, dh, scale,
                                        softcap, _ptr(opart), _ptr(lse), _stream()), "xattn_splitkv_seg")
    return opart, lse


def xattn_splitkv(q, k, v, kmask, Hq: int, Hkv: int, dh: int, scale: float, softcap: float, splits: int,
                  opart=None, lse=None, impl: str = "auto"):
    """q [T, Hq*dh]; k,v [N, *] views with row stride ld -> (opart [splits,T,Hq,dh] f32, lse [splits,T,Hq] f32)."""
    L = _lib.load()
    T = q.shape[0]
    N = k.shape[0]
    assert q.dtype == BF16 and k.dtype == BF16 and v.dtype == BF16
    assert k.stride(0) == v.stride(0) or N == 0
    if opart is None:
        opart = torch.empty(splits, T, Hq, dh, device=q.device, dtype=torch.float32)
        lse = torch.empty(splits, T, Hq, device=q.device, dtype=torch.float32)
    fn = L.vidi_xattn_splitkv_mma if impl == "mma" else L.vidi_xattn_splitkv
    _lib.check(fn(_ptr(q), _rowmajor(q), _ptr(k), _ptr(v), k.stride(0) if N else 8, _ptr(kmask), T, N,
                  Hq, Hkv, dh, splits, scale, softcap, _ptr(opart), _ptr(lse), _stream()), "xattn_splitkv")
    return opart, lse


def xattn_merge(opart, lse, out, gate: float = 1.0, accumulate: bool = False, P=None, splits_per_rank=None,
                rank_stride_o: int = 0, rank_stride_l: int = 0, rows=None, dh=None):
    """Merge partials into out fp32 [rows, dh] (+=).  Default: opart [P, rows, dh], lse [P, rows] contiguous.
    With explicit P / splits_per_rank / rank strides the partials may sit in an all-gathered flat buffer."""
    L = _lib.load()
    if P is None:
        P = opart.shape[0]
        dh = opart.shape[-1]
        rows = lse.numel() // P
        splits_per_rank = P
        assert opart.is_contiguous() and lse.is_contiguous()
    assert out.dtype == torch.float32 and out.is_contiguous()
    _lib.check(L.vidi_xattn_merge(_ptr(opart), _ptr(lse), P, splits_per_rank, rank_stride_o, rank_stride_l, rows, dh, gate,
                                  int(accumulate), _ptr(out), _stream()), "xattn_merge")
    return out


def text_qk_prep(qkv, q_rope, kv_out, Hq: int, Hkv: int, dh: int, inv_freq, pos0: int = 0):
    """qkv [Tq, (Hq+2Hkv)*dh] -> q_rope [Tq, Hq*dh] = RoPE(q); kv_out [Tq, 2*Hkv*dh] = RoPE(k) | v."""
    L = _lib.load()
    Tq = qkv.shape[0]
    _lib.check(L.vidi_text_qk_prep(_ptr(qkv), _rowmajor(qkv), _ptr(q_rope), _rowmajor(q_rope), _ptr(kv_out), _rowmajor(kv_out),
                                   Tq, Hq, Hkv, dh, _ptr(inv_freq), pos0, _stream()), "text_qk_prep")
    return q_rope, kv_out


def xattn_merge2(srcs, att, out_bf16, rows: int, dh: int):
    """srcs: up to two (O, LSE, P, splits_per_rank, rank_stride_o, rank_stride_l, gate); out = bf16(att + sum merged)."""
    L = _lib.load()
    assert len(srcs) <= 2 and att.dtype == torch.float32 and out_bf16.dtype == BF16
    pad = (None, None, 0, 1, 0, 0, 0.0)
    s0 = srcs[0] if len(srcs) > 0 else pad
    s1 = srcs[1] if len(srcs) > 1 else pad
    _lib.check(L.vidi_xattn_merge2(_ptr(s0[0]), _ptr(s0[1]), s0[2], s0[3], s0[4], s0[5], s0[6], _ptr(s1[0]), _ptr(s1[1]), s1[2],
                                   s1[3], s1[4], s1[5], s1[6], len(srcs), _ptr(att), rows, dh, _ptr(out_bf16), _stream()),
               "xattn_merge2")
    return out_bf16


def xattn_premerge(srcs, rows: int, dh: int, out: torch.Tensor):
    """srcs: [(O [P,rows,dh] f32, LSE [P,rows] f32, P)] per stream -> out fp32 flat [per stream: O [rows,dh] | LSE [rows]]: this rank's
    key splits reduced to one partial per stream (what crosses ranks; csrc/xchg.cu)."""
    L = _lib.load()
    assert 1 <= len(srcs) <= 2 and out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= len(srcs) * rows * (dh + 1)
    s0 = srcs[0]
    s1 = srcs[1] if len(srcs) > 1 else (None, None, 0)
    base = (C.c_void_p * 1)(out.data_ptr())
    _lib.check(L.vidi_xattn_premerge_push(_ptr(s0[0]), _ptr(s0[1]), s0[2], _ptr(s1[0]), _ptr(s1[1]), s1[2], len(srcs), rows, dh, base,
                                          None, 1, 0, 0, None, _stream()), "xattn_premerge_push")
    return out


def xchg_push(xchg, srcs, rows: int, dh: int) -> int:
    """send side of the peer-memory exchange (exchange.PartialExchange.push) on the current stream"""
    return xchg.push([(_ptr(o), _ptr(l), p) for o, l, p in srcs], rows, dh, _stream())


def xchg_merge(xchg, gates, att, out_bf16, rows: int, dh: int):
    """receive side: waits for every rank's partial of the current exchange, then out = bf16(att + sum_s gate_s * merge_s)"""
    assert att.dtype == torch.float32 and out_bf16.dtype == BF16 and att.is_cuda
    xchg.merge(gates, att, out_bf16, rows, dh, _stream())
    return out_bf16


def rope_inplace(x, col_off: int, heads: int, dh: int, inv_freq, pos0: int = 0):
    L = _lib.load()
    T = x.shape[0]
    _lib.check(L.vidi_rope_inplace(_ptr(x), _rowmajor(x), col_off, T, heads, dh, _ptr(inv_freq), pos0, _stream()),
               "rope_inplace")
    return x


def attn_text(q, k, v, pos0: int, Hq: int, Hkv: int, dh: int, scale: float, softcap: float, window: int, out=None):
    """q [Tq, Hq*dh] view, k/v [Tk, Hkv*dh] views -> fp32 [Tq, Hq*dh]."""
    L = _lib.load()
    Tq, Tk = q.shape[0], k.shape[0]
    if out is None:
        out = torch.empty(Tq, Hq * dh, device=q.device, dtype=torch.float32)
    _lib.check(L.vidi_attn_text(_ptr(q), _rowmajor(q), _ptr(k), _ptr(v), _rowmajor(k), Tq, Tk, pos0, Hq, Hkv, dh, scale,
                                softcap, window, _ptr(out), _stream()), "attn_text")
    return out


def launch_count() -> int:
    return int(_lib.load().vidi_launch_count())


def reset_launch_count() -> None:
    _lib.load().vidi_reset_launch_count()


# ------------------------------------------------------------------------------------------------------
# bench instrumentation for the non-GEMM ops: when PROFILE_OPS is a dict, every wrapped op accumulates
# (calls, [event pairs]) under its name.  Zero overhead when PROFILE_OPS is None.
# ------------------------------------------------------------------------------------------------------
PROFILE_OPS = None


def _instrument(fn):
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        prof = PROFILE_OPS
        if prof is None:
            return fn(*a, **k)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        prof.setdefault(fn.__name__, []).append((e0, e1))
        return r
    return wrapped


for _name in ("rmsnorm", "residual_norm", "layernorm", "mm_finish", "rmsnorm_f32", "patch_im2col", "whisper_im2col1",
              "whisper_im2col2", "pool_s2d", "conv_window_gather", "bilinear_ac", "embed_gather", "sinusoid_split", "split3",
              "cast_bf16", "attn_dense", "xattn_splitkv", "xattn_merge", "rope_inplace", "attn_text", "text_qk_prep", "xattn_merge2",
              "xattn_premerge", "xchg_push", "xchg_merge", "xattn_splitkv_seg"):
    globals()[_name] = _instrument(globals()[_name])


_RESAMPLE_TABLES = {}


def _resample_tables(in_size: int, out_size: int, device):
    """Pillow tap tables (preprocess.pil_bicubic_coeffs) as int32 device tensors, cached per (in, out, device)."""
    key = (in_size, out_size, str(device))
    if key not in _RESAMPLE_TABLES:
        from .preprocess import pil_bicubic_coeffs
        xmin, kk = pil_bicubic_coeffs(in_size, out_size)
        _RESAMPLE_TABLES[key] = (xmin.to(device=device, dtype=torch.int32).contiguous(), kk.to(device=device, dtype=torch.int32).contiguous())
    return _RESAMPLE_TABLES[key]


def resize_frames_u8(frames: torch.Tensor, size: int, rescale: float = 1.0 / 255.0, mean: float = 0.5, std: float = 0.5) -> torch.Tensor:
    """frames uint8 [F,H,W,3] (decoded RGB, on the GPU) -> bf16 [F,3,size,size] = ((PIL_bicubic(frames) * rescale) - mean) / std:
    process_images' 'resize' branch + SiglipImageProcessor (img_utils.py:181-187), resampling bit-exact with Pillow."""
    L = _lib.load()
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3 and frames.is_cuda and frames.is_contiguous()
    F, H, W, _ = frames.shape
    out = torch.empty(F, 3, size, size, device=frames.device, dtype=BF16)
    if F == 0:
        return out
    mid = frames
    if W != size:                                   # Pillow skips a pass whose size does not change
        xmin, kk = _resample_tables(W, size, frames.device)
        mid = torch.empty(F, H, size, 3, device=frames.device, dtype=torch.uint8)
        _lib.check(L.vidi_resample_u8(_ptr(frames), _ptr(mid), F * H, W, size, 3, _ptr(xmin), _ptr(kk), kk.shape[1], _stream()),
                   "resample_u8")
    ymin, kk = _resample_tables(H, size, frames.device)  # in == out gives the identity taps exactly (weight 1 << 22 on the centre)
    _lib.check(L.vidi_resample_u8_to_chw_bf16(_ptr(mid), _ptr(out), F, H, size, size, _ptr(ymin), _ptr(kk), kk.shape[1],
                                              float(rescale), float(mean), float(std), _stream()), "resample_u8_to_chw_bf16")
    return out



_LOGMEL_TABLES = {}


def _logmel_tables(mels: int, device):
    """hann window, split DFT matrix [408, 1200] (rows: cos 0..200 | -sin 0..200 | 0) and split mel matrix [mels, 624], built in fp64"""
    key = (mels, str(device))
    if key not in _LOGMEL_TABLES:
        import math
        from .preprocess import mel_filter_bank
        win = torch.hann_window(400, periodic=True, dtype=torch.float64)
        ang = torch.arange(201, dtype=torch.float64)[:, None] * torch.arange(400, dtype=torch.float64)[None, :] * (2 * math.pi / 400)
        dft = torch.zeros(408, 400, dtype=torch.float64)
        dft[:201], dft[201:402] = torch.cos(ang), -torch.sin(ang)
        fb = torch.zeros(mels, 208, dtype=torch.float64)
        fb[:, :201] = mel_filter_bank(201, mels, 16000).t()
        _LOGMEL_TABLES[key] = (win.float().to(device), split3(dft.float().to(device).contiguous(), 1),
                               split3(fb.float().to(device).contiguous(), 1))
    return _LOGMEL_TABLES[key]


def log_mel(chunks: torch.Tensor, mels: int = 128, group: int = 16) -> torch.Tensor:
    """chunks fp32 [C, 480000] (zero-padded 30-s windows of 16 kHz audio, on the GPU) -> bf16 [C, mels, 3000] Whisper log-mel features
    (WhisperFeatureExtractor semantics, vid_utils.py:52-63); DFT and mel projection on the tensor cores in 3-term split-bf16 form."""
    L = _lib.load()
    assert chunks.dtype == torch.float32 and chunks.dim() == 2 and chunks.shape[1] == 480000 and chunks.is_cuda and chunks.is_contiguous()
    C = chunks.shape[0]
    out = torch.empty(C, mels, 3000, device=chunks.device, dtype=BF16)
    if C == 0:
        return out
    win, wdft, wmel = _logmel_tables(mels, chunks.device)
    for c0 in range(0, C, group):                                   # bounds the [rows, 1200] operand (16 chunks = 115 MB)
        c1 = min(C, c0 + group)
        rows = (c1 - c0) * 3001
        a1 = torch.empty(rows, 1200, device=chunks.device, dtype=BF16)
        _lib.check(L.vidi_logmel_frames(_ptr(chunks[c0:c1]), _ptr(win), _ptr(a1), c1 - c0, 480000, _stream()), "logmel_frames")
        y = gemm(a1, wdft, out_fp32=True, tag="logmel", alg_k=400)
        a2 = torch.empty(rows, 624, device=chunks.device, dtype=BF16)
        _lib.check(L.vidi_logmel_power(_ptr(y), y.stride(0), _ptr(a2), rows, _stream()), "logmel_power")
        m = gemm(a2, wmel, out_fp32=True, tag="logmel", alg_k=201)
        cmax = torch.empty(c1 - c0, device=chunks.device, dtype=torch.float32)
        _lib.check(L.vidi_logmel_finish(_ptr(m), c1 - c0, mels, _ptr(cmax), _ptr(out[c0:c1]), _stream()), "logmel_finish")
    return out

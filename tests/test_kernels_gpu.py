"""Per-kernel parity of the sm_100a library (through the C ABI) against plain PyTorch fp32 / the oracle."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def dev():
    return torch.device("cuda:0")


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda"); g.manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g, device="cuda") * scale)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 256), (128, 64, 64, 64), (256, 128, 128, 128),
    (300, 520, 288, 256), (300, 520, 288, 128), (77, 72, 1152, 64),
    (1000, 1152, 640, 256), (2048, 4096, 3584, 256), (4096, 3584, 2048, 128),
    (33, 1024, 512, 256), (5000, 4304, 1152, 256), (1000, 1152, 1152, 192), (300, 3456, 288, 192), (129, 384, 72, 192),
])
def test_gemm_plain(M, N, K, bn):
    from vidi_b200 import ops
    a = rnd(M, K, seed=1).to(BF); w = rnd(N, K, scale=0.05, seed=2).to(BF)
    out = ops.gemm(a, w, block_n=bn)
    ref = a.float() @ w.float().t()
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 6e-3, rel_err(out, ref)
    assert float((out.float() - ref).abs().max()) < 0.05 * float(ref.abs().max()) + 1e-2


@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 64, 256), (256, 128, 128, 128), (300, 520, 288, 256), (1000, 1152, 640, 128),
                                      (4096, 4096, 3584, 256), (5000, 4304, 1152, 256), (129, 72, 64, 128), (46656, 1152, 1152, 128), (3000, 1152, 4304, 192),
                                      (700, 3456, 1152, 192), (3000, 1152, 4304, 256), (700, 3456, 1152, 256), (1030, 1296, 320, 256),
                                      (513, 264, 128, 256)])
def test_gemm_2cta(M, N, K, bn):
    """CTA-pair (cta_group::2) kernel: same contract as the 1-CTA kernel.  N not a multiple of block_n exercises the ragged last
    column tile, which runs a narrower MMA (n_eff = roundup(N - n0, 16): 128 for N = 1152 / 3456, 208 for 4304, 16 for 1296, 8 -> 16 for
    264) with each CTA of the pair supplying n_eff / 2 rows of W."""
    from vidi_b200 import ops
    a = rnd(M, K, seed=51).to(BF); w = rnd(N, K, scale=0.05, seed=52).to(BF)
    bias = rnd(N, seed=53).float(); res = rnd(M, N, seed=54).to(BF)
    out = ops.gemm(a, w, bias=bias, residual=res, act=2, block_n=bn, cta2=True)
    ref = F.gelu(a.float() @ w.float().t() + bias, approximate="tanh") + res.float()
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 6e-3, rel_err(out, ref)


def test_gemm_2cta_glu():
    from vidi_b200 import ops
    from vidi_b200.weights import pack_glu
    M, I, K = 1300, 1024, 512
    a = rnd(M, K, seed=55).to(BF)
    wg = rnd(I, K, scale=0.05, seed=56).to(BF); wu = rnd(I, K, scale=0.05, seed=57).to(BF)
    out = ops.gemm(a, pack_glu(wg, wu, 256), glu=1, block_n=256, cta2=True)
    ref = F.gelu(a.float() @ wg.float().t(), approximate="tanh") * (a.float() @ wu.float().t())
    assert out.shape == (M, I) and rel_err(out, ref) < 8e-3


@pytest.mark.parametrize("M,D,N,act", [(1000, 1152, 3456, 0), (700, 1152, 4304, 2), (300, 288, 520, 1), (129, 1280, 5120, 1),
                                       (2600, 256, 768, 0)])
def test_gemm_layernorm_fold(M, D, N, act):
    """Pre-LN block without a LayerNorm kernel: GEMM 1 (+bias +residual) writes x and its row statistics, GEMM 2 consumes the raw x
    with gamma/beta folded into its weights and applies mean / rstd in the epilogue  ==  Linear(LayerNorm(x))."""
    from vidi_b200 import ops
    from vidi_b200.weights import fold_layernorm
    K0 = 320
    a = rnd(M, K0, seed=61).to(BF); w0 = rnd(D, K0, scale=0.06, seed=62).to(BF)
    b0 = rnd(D, seed=63).float(); res = (rnd(M, D, seed=64) * 2 + 0.7).to(BF)          # non-zero row means
    st = torch.full((M, ops.ln_stats_parts(D), 2), float("nan"), device="cuda")
    x = ops.gemm_ln(a, w0, bias=b0, residual=res, stats=st)
    x_ref = (a.float() @ w0.float().t() + b0 + res.float())
    assert rel_err(x, x_ref) < 6e-3
    xs = x.float()
    assert torch.allclose(st[..., 0].sum(1), xs.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[..., 1].sum(1), (xs * xs).sum(1), rtol=1e-4, atol=1e-2)
    gamma = (1 + rnd(D, scale=0.2, seed=65)).float(); beta = rnd(D, scale=0.3, seed=66).float()
    w = rnd(N, D, scale=0.05, seed=67).to(BF); b = rnd(N, seed=68).float()
    wp, cs, bp = fold_layernorm(w, b, gamma, beta)
    st2 = torch.empty(M, ops.ln_stats_parts(N), 2, device="cuda")
    y = ops.gemm_ln(x, wp, bias=bp, act=act, ln=(st, cs, 1e-6), stats=st2)
    h = F.layer_norm(xs, (D,), gamma, beta, 1e-6)
    ref = h @ w.float().t() + b
    if act == 1: ref = F.gelu(ref)
    elif act == 2: ref = F.gelu(ref, approximate="tanh")
    assert rel_err(y, ref) < 8e-3, rel_err(y, ref)
    assert torch.allclose(st2[..., 0].sum(1), y.float().sum(1), rtol=1e-3, atol=5e-2)


@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
def test_gemm_epilogues(act):
    from vidi_b200 import ops
    M, N, K = 515, 712, 264
    a = rnd(M, K, seed=3).to(BF); w = rnd(N, K, scale=0.1, seed=4).to(BF)
    bias = rnd(N, seed=5).float()
    res = rnd(M, N, seed=6).to(BF)
    out = ops.gemm(a, w, bias=bias, residual=res, act=act, act_param=3.0)
    x = a.float() @ w.float().t() + bias
    if act == 1: x = F.gelu(x)
    elif act == 2: x = F.gelu(x, approximate="tanh")
    elif act == 3: x = 3.0 * torch.tanh(x / 3.0)
    elif act == 4: x = F.silu(x)
    ref = x + res.float()
    assert rel_err(out, ref) < 6e-3
    # fp32 output, pos-emb style residual (row % res_mod)
    pos = rnd(100, N, seed=7).to(BF)
    out32 = ops.gemm(a, w, bias=bias, residual=pos, res_mod=100, out_fp32=True)
    ref32 = a.float() @ w.float().t() + bias + pos.float()[torch.arange(M, device="cuda") % 100]
    assert out32.dtype == torch.float32 and rel_err(out32, ref32) < 2e-3


@pytest.mark.parametrize("glu", [1, 2])
def test_gemm_glu(glu):
    from vidi_b200 import ops
    from vidi_b200.weights import pack_glu
    M, I, K = 700, 1024, 512
    a = rnd(M, K, seed=8).to(BF)
    wg = rnd(I, K, scale=0.05, seed=9).to(BF); wu = rnd(I, K, scale=0.05, seed=10).to(BF)
    wp = pack_glu(wg, wu, 256)
    out = ops.gemm(a, wp, glu=glu, block_n=256)
    g = a.float() @ wg.float().t(); u = a.float() @ wu.float().t()
    ref = (F.gelu(g, approximate="tanh") if glu == 1 else F.silu(g)) * u
    assert out.shape == (M, I) and rel_err(out, ref) < 8e-3


def test_gemm_strided_views():
    from vidi_b200 import ops
    big = rnd(400, 1024, seed=11).to(BF)
    a = big[:, 256:768]                      # lda = 1024, K = 512
    w = rnd(384, 512, scale=0.05, seed=12).to(BF)
    outbuf = torch.zeros(400, 1024, device="cuda", dtype=BF)
    ops.gemm(a, w, out=outbuf[:, 128:512])
    assert rel_err(outbuf[:, 128:512], a.float() @ w.float().t()) < 6e-3
    assert float(outbuf[:, :128].abs().max()) == 0 and float(outbuf[:, 512:].abs().max()) == 0


# ------------------------------------------------------------------ norms
@pytest.mark.parametrize("D", [3584, 512, 1152, 4096])
def test_rmsnorm(D):
    from vidi_b200 import ops
    x = rnd(1001, D, scale=3.0, seed=13).to(BF); w = rnd(D, scale=0.1, seed=14).to(BF)
    for add_one in (True, False):
        y = ops.rmsnorm(x, w, 1e-6, add_one)
        xf = x.float(); xh = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)
        ref = xh * ((1 + w.float()) if add_one else w.float())
        assert rel_err(y, ref) < 4e-3


@pytest.mark.parametrize("post_mode,next_add_one", [(1, True), (0, False)])
def test_residual_norm(post_mode, next_add_one):
    from vidi_b200 import ops
    D = 3584
    x = rnd(777, D, seed=15).to(BF); y = rnd(777, D, scale=2.0, seed=16).to(BF)
    wp = rnd(D, scale=0.1, seed=17).to(BF); wn = rnd(D, scale=0.1, seed=18).to(BF)
    x0 = x.clone(); h = torch.empty_like(x)
    ops.residual_norm(x, y, wp, wn, h, 1e-6, post_mode, next_add_one)
    yf = y.float()
    if post_mode:
        yf = (yf * torch.rsqrt(yf.pow(2).mean(-1, keepdim=True) + 1e-6) * (1 + wp.float())).to(BF).float()
    xr = (x0.float() + yf).to(BF)
    assert rel_err(x, xr) < 4e-3
    xf = xr.float(); hr = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * ((1 + wn.float()) if next_add_one else wn.float())
    assert rel_err(h, hr) < 5e-3


@pytest.mark.parametrize("D", [1152, 1280, 288])
def test_layernorm(D):
    from vidi_b200 import ops
    x = (rnd(999, D, scale=2.0, seed=19) + 0.5).to(BF)
    w = (1 + rnd(D, scale=0.1, seed=20)).float(); b = rnd(D, scale=0.1, seed=21).float()
    y = ops.layernorm(x, w, b, 1e-6)
    assert rel_err(y, F.layer_norm(x.float(), (D,), w, b, 1e-6)) < 4e-3


def test_mm_finish_matches_oracle_algebra():
    from vidi_b200 import ops
    D, Fr, hp, wp = 512, 5, 14, 14
    rows = Fr * hp * wp
    proj = rnd(rows, D, scale=2.0, seed=22).to(BF)
    w_mod = (1 + rnd(D, scale=0.1, seed=23)).to(BF); w_llm = ((1 + rnd(D, scale=0.1, seed=24)) * 0.029).to(BF)
    th, tw, tt = rnd(hp, D, seed=25).float().contiguous(), rnd(wp, D, seed=26).float().contiguous(), rnd(Fr + 3, D, seed=27).float().contiguous()
    nrm = float(torch.tensor(D ** 0.5, dtype=BF).float())
    out, mask = ops.mm_finish(proj, w_mod, w_llm, [th, tw, tt], [wp, 1, hp * wp], [hp, wp, 1 << 30], [0, 0, 2], 0, True, nrm, 1e-5)
    xf = proj.float(); x = w_mod.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5))
    x = x.view(Fr, hp, wp, D) + th[None, :, None] + tw[None, None, :] + tt[2:2 + Fr][:, None, None]
    x = x.view(rows, D)
    ref = w_llm.float() * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5)) * nrm
    assert bool(mask.all()) and rel_err(out, ref) < 8e-3
    out0, mask0 = ops.mm_finish(proj, w_mod, w_llm, [th, tw, tt], [wp, 1, hp * wp], [hp, wp, 1 << 30], [0, 0, 2], 0, False, nrm, 1e-5)
    assert not bool(mask0.any()) and float(out0.abs().max()) == 0.0


# ------------------------------------------------------------------ layout kernels vs the oracle
def test_patch_im2col_and_conv():
    from vidi_b200 import ops
    img = rnd(3, 3, 378, 378, seed=28).clamp(-1, 1).to(BF)
    w = rnd(64, 3, 14, 14, scale=0.05, seed=29).to(BF)
    A = ops.patch_im2col(img, 14, 640)
    wp = torch.zeros(64, 640, device="cuda", dtype=BF); wp[:, :588] = w.reshape(64, 588)
    out = ops.gemm(A, wp)
    ref = F.conv2d(img.float(), w.float(), stride=14).flatten(2).transpose(1, 2).reshape(-1, 64)
    assert rel_err(out, ref) < 6e-3


@pytest.mark.parametrize("hw", [(28, 28), (20, 20), (10, 10), (16, 12)])
def test_pool_s2d(hw):
    from vidi_b200 import ops
    from oracle import vidi15_ref as R
    Fr, side, d = 3, 27, 64
    P = rnd(Fr, side * side, d, seed=30).to(BF)
    X = ops.pool_s2d(P, Fr, side, hw[0], hw[1], 2)
    feats = P.float().cpu().reshape(Fr, side, side, d).permute(0, 3, 1, 2)
    ref = R.conv2d_pool(feats, hw, 2).permute(0, 2, 3, 1)                       # [F,h',w', c*4+q]
    ref = ref.reshape(Fr, hw[0] // 2, hw[1] // 2, d, 4).permute(0, 1, 2, 4, 3).reshape(-1, 4 * d)   # -> q*d+c
    assert rel_err(X.cpu(), ref) < 5e-3


@pytest.mark.parametrize("s_out", [16, 4, 9])
def test_vidi7b_conv_pool(s_out):
    """Vidi-7B Conv2DPool: window gather + GEMM + bilinear(align_corners=True) vs the oracle (pool.py:6-26)."""
    import math
    from vidi_b200 import ops
    from oracle import vidi7b_ref as R7
    Fr, side, d = 2, 27, 64
    k = math.ceil(side / s_out)
    P = rnd(Fr, side * side, d, seed=40).to(BF)
    w = rnd(d, d, k, k, scale=0.05, seed=41).to(BF)
    A = ops.conv_window_gather(P, Fr, side, k)
    Y = ops.gemm(A, w.permute(0, 2, 3, 1).reshape(d, -1).contiguous())
    X = ops.bilinear_ac(Y, Fr, side - k + 1, s_out)
    feats = P.float().cpu().reshape(Fr, side, side, d).permute(0, 3, 1, 2)
    ref = R7.conv2d_pool_7b(feats, w.float().cpu(), s_out).permute(0, 2, 3, 1).reshape(-1, d)
    assert rel_err(X.cpu(), ref) < 8e-3


def test_whisper_im2col():
    from vidi_b200 import ops
    Cn, mels, T, d = 2, 128, 3000, 64
    mel = rnd(Cn, mels, T, seed=31).to(BF)
    w1 = rnd(d, mels, 3, scale=0.05, seed=32).to(BF)
    A1 = ops.whisper_im2col1(mel)
    y1 = ops.gemm(A1, w1.permute(0, 2, 1).reshape(d, 3 * mels).contiguous())
    ref1 = F.conv1d(mel.float(), w1.float(), padding=1).permute(0, 2, 1).reshape(-1, d)
    assert rel_err(y1, ref1) < 6e-3
    w2 = rnd(d, d, 3, scale=0.05, seed=33).to(BF)
    A2 = ops.whisper_im2col2(y1, Cn, T)
    y2 = ops.gemm(A2, w2.permute(0, 2, 1).reshape(d, 3 * d).contiguous())
    ref2 = F.conv1d(y1.float().view(Cn, T, d).permute(0, 2, 1), w2.float(), stride=2, padding=1).permute(0, 2, 1).reshape(-1, d)
    assert rel_err(y2, ref2) < 6e-3


def test_embed_gather_and_pos_split():
    from vidi_b200 import ops
    from oracle import vidi15_ref as R
    E = rnd(1000, 512, scale=0.02, seed=34).to(BF)
    ids = torch.tensor([2, 5, 999, 0, 17], device="cuda")
    out = ops.embed_gather(ids, E, 22.625)
    assert torch.equal(out, (E[ids].float() * 22.625).to(BF))
    # split-precision positional MLP ~ fp32 accuracy
    D, l, N = 512, 37, 10000
    div = torch.exp(torch.arange(0, D, 2, dtype=torch.float) * -(math.log(10000.0) / D)).cuda()
    A = ops.sinusoid_split(div, l, 0, l, N, D)
    pe_ref = R.sinusoid(torch.arange(l, dtype=torch.float) / (l - 1) * (N - 1), D)
    pe = (A[:, :D].float() + A[:, 2 * D:].float()).cpu()
    assert float((pe - pe_ref).abs().max()) < 2e-3        # sin/cos of large arguments: device vs host libm
    W = rnd(D, D, scale=0.03, seed=35).float()
    Wp = ops.split3(W.contiguous(), 1)
    y = ops.gemm(A, Wp, out_fp32=True)
    ref = (A[:, :D].float() + A[:, 2 * D:].float()) @ W.t()
    assert rel_err(y, ref) < 3e-5


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("impl", ["auto", "v1", "v2", "mma"])
@pytest.mark.parametrize("B,S,H,dh", [(3, 729, 4, 72), (2, 1500, 4, 64), (1, 100, 2, 72), (2, 64, 2, 64), (5, 729, 16, 72),
                                      (1, 129, 1, 72), (3, 257, 3, 64)])
def test_attn_dense(B, S, H, dh, impl):
    from vidi_b200 import ops
    d = H * dh
    qkv = rnd(B * S, 3 * d, seed=36).to(BF)
    out = ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, impl=impl)
    q, k, v = [t.float().view(B, S, H, dh).transpose(1, 2) for t in qkv.split(d, dim=1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v).transpose(1, 2).reshape(B * S, d)
    assert rel_err(out, ref) < 8e-3


@pytest.mark.parametrize("S,H,dh,mode", [(729, 4, 72, "rising"), (1500, 4, 64, "rising"), (729, 4, 72, "falling"),
                                         (700, 2, 64, "spiky")])
@pytest.mark.parametrize("impl", ["auto", "v2"])
def test_attn_dense_reference_moves(S, H, dh, mode, impl):
    """Score ranges that drift by far more than 2^8 between key tiles: exercises the lazy re-referencing of the
    ping-pong softmax (attn2_sm100.cu) both ways (reference must move / must not move)."""
    from vidi_b200 import ops
    B, d = 3, H * dh
    g = torch.Generator(device="cuda").manual_seed(91)
    qkv = torch.randn(B * S, 3 * d, device="cuda", generator=g)
    pos = torch.arange(S, device="cuda", dtype=torch.float32).repeat(B)
    if mode == "rising":
        gain = 0.2 + pos / 64.0                      # later keys much larger
    elif mode == "falling":
        gain = 0.2 + (S - pos) / 64.0
    else:
        gain = torch.where((pos.long() % 197) == 190, 30.0, 0.5)   # isolated huge keys in late tiles
    qkv[:, d:2 * d] *= gain[:, None]
    qkv[:, :d] = qkv[:, :d].abs()                    # positive q·k drift: |q|·k with k scaled → wide logit range
    qkv = qkv.to(BF)
    out = ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, impl=impl)
    q, k, v = [t.float().view(B, S, H, dh).transpose(1, 2) for t in qkv.split(d, dim=1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v).transpose(1, 2).reshape(B * S, d)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("impl", ["auto", "mma"])
@pytest.mark.parametrize("T,N,Hq,Hkv,dh,cap,splits", [
    (32, 5000, 16, 8, 256, 50.0, 7), (12, 300, 4, 2, 256, 50.0, 3), (40, 2048, 32, 8, 128, 0.0, 4),
    (33, 1000, 16, 8, 256, 50.0, 1), (5, 31, 4, 2, 256, 50.0, 2), (64, 4097, 16, 8, 256, 50.0, 5), (70, 700, 16, 8, 256, 50.0, 3),
    (32, 126000, 16, 8, 256, 50.0, 37),
])
def test_xattn_splitkv_and_merge(T, N, Hq, Hkv, dh, cap, splits, impl):
    from vidi_b200 import ops
    q = rnd(T, Hq * dh, seed=37).to(BF)
    kv = rnd(N, 2 * Hkv * dh, seed=38).to(BF)
    k, v = kv[:, :Hkv * dh], kv[:, Hkv * dh:]
    mask = torch.ones(N, device="cuda", dtype=torch.uint8); mask[N // 3: N // 3 + 5] = 0
    scale = dh ** -0.5
    op, lse = ops.xattn_splitkv(q, k, v, mask, Hq, Hkv, dh, scale, cap, splits, impl=impl)
    out = torch.zeros(T * Hq, dh, device="cuda")
    ops.xattn_merge(op, lse, out)
    qh = q.float().view(T, Hq, dh).transpose(0, 1)
    kh = k.float().reshape(N, Hkv, dh).transpose(0, 1).repeat_interleave(Hq // Hkv, 0)
    vh = v.float().reshape(N, Hkv, dh).transpose(0, 1).repeat_interleave(Hq // Hkv, 0)
    s = qh @ kh.transpose(-1, -2) * scale
    if cap > 0: s = cap * torch.tanh(s / cap)
    s = s.masked_fill(mask[None, None, :] == 0, float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(0, 1).reshape(T * Hq, dh)
    assert rel_err(out, ref) < 8e-3
    # accumulate + gate
    ops.xattn_merge(op, lse, out, gate=0.5, accumulate=True)
    assert rel_err(out, 1.5 * ref) < 8e-3
    # merging the partials in two groups (== two ranks) then merging again must agree
    if splits >= 2:
        lse_all = torch.logsumexp(lse.view(splits, -1), 0)
        sref = torch.logsumexp(s, -1).transpose(0, 1).reshape(-1)
        assert float((lse_all - sref).abs().max()) < 2e-2


def test_rope_and_attn_text():
    from vidi_b200 import ops
    from oracle import vidi15_ref as R
    T, Hq, Hkv, dh = 37, 4, 2, 256
    qd, kd = Hq * dh, Hkv * dh
    qkv = rnd(T, qd + 2 * kd, seed=39).to(BF)
    ref_in = qkv.clone()
    inv = (1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.float) / dh))).cuda()
    ops.rope_inplace(qkv, 0, Hq, dh, inv)
    ops.rope_inplace(qkv, qd, Hkv, dh, inv)
    cos, sin = R.rope_cos_sin(T, dh, 10000.0)
    qr = R.apply_rope(ref_in[:, :qd].float().cpu().view(T, Hq, dh).transpose(0, 1), cos, sin)
    kr = R.apply_rope(ref_in[:, qd:qd + kd].float().cpu().view(T, Hkv, dh).transpose(0, 1), cos, sin)
    assert rel_err(qkv[:, :qd].cpu().view(T, Hq, dh).transpose(0, 1), qr) < 6e-3
    for window in (0, 8):
        out = ops.attn_text(qkv[:, :qd], qkv[:, qd:qd + kd], qkv[:, qd + kd:], 0, Hq, Hkv, dh, 1 / 16, 50.0, window)
        i = torch.arange(T)
        allowed = i[None, :] <= i[:, None]
        if window: allowed &= (i[:, None] - i[None, :] < window)
        bias = torch.zeros(T, T).masked_fill(~allowed, float("-inf"))
        vh = ref_in[:, qd + kd:].float().cpu().view(T, Hkv, dh).transpose(0, 1)
        ref = R.attend(qkv[:, :qd].float().cpu().view(T, Hq, dh).transpose(0, 1),
                       qkv[:, qd:qd + kd].float().cpu().view(T, Hkv, dh).transpose(0, 1), vh, 1 / 16, 50.0, bias)
        assert rel_err(out.cpu(), ref) < 5e-3


def test_text_qk_prep_and_merge2_match_unfused_path():
    """The fused text helpers must equal the individual kernels they replace (clone+rope+copy, merge x2 + cast)."""
    from vidi_b200 import ops
    T, Hq, Hkv, dh = 19, 4, 2, 256
    qd, kd = Hq * dh, Hkv * dh
    qkv = rnd(T, qd + 2 * kd, seed=60).to(BF)
    inv = (1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.float) / dh))).cuda()
    q_ref = qkv[:, :qd].clone(); kv_ref = qkv[:, qd:].clone()
    ops.rope_inplace(q_ref, 0, Hq, dh, inv, 7); ops.rope_inplace(kv_ref, 0, Hkv, dh, inv, 7)
    q_out = torch.empty(T, qd, device="cuda", dtype=BF); kv_out = torch.empty(T, 2 * kd, device="cuda", dtype=BF)
    ops.text_qk_prep(qkv, q_out, kv_out, Hq, Hkv, dh, inv, 7)
    assert torch.equal(q_out, q_ref) and torch.equal(kv_out, kv_ref)
    rows = T * Hq
    att = rnd(rows, dh, seed=61).float().contiguous()
    srcs, ref = [], att.clone()
    for i, (P, gate) in enumerate([(5, 1.0), (3, 0.5)]):
        O = rnd(P, rows, dh, seed=62 + i).float().contiguous(); Ls = rnd(P, rows, seed=64 + i).float().contiguous()
        Ls[1, ::3] = float("-inf")
        ops.xattn_merge(O, Ls, ref, gate=gate, accumulate=True)
        srcs.append((O, Ls, P, P, 0, 0, gate))
    out = torch.empty(rows, dh, device="cuda", dtype=BF)
    ops.xattn_merge2(srcs, att, out, rows, dh)
    assert rel_err(out, ref) < 4e-3
    out1 = torch.empty(rows, dh, device="cuda", dtype=BF)
    ops.xattn_merge2([], att, out1, rows, dh)
    assert torch.equal(out1, att.to(BF))


def _merge_ref(parts):
    """fp64 reference: merge a list of (O [rows, dh], LSE [rows]) partials -> (O, LSE)"""
    lse = torch.stack([l for _, l in parts]).double()
    Lm = lse.max(0).values
    Lm = torch.where(torch.isinf(Lm), torch.zeros_like(Lm), Lm)
    w = torch.exp(lse - Lm)
    den = w.sum(0)
    O = sum(w[i][:, None] * parts[i][0].double() for i in range(len(parts))) / den.clamp_min(1e-300)[:, None]
    O = torch.where(den[:, None] > 0, O, torch.zeros_like(O))
    return O, torch.where(den > 0, Lm + torch.log(den), torch.full_like(den, float("-inf")))


@pytest.mark.parametrize("world,dh", [(2, 256), (3, 128), (8, 256)])
def test_premerge_then_rank_strided_merge2(world, dh):
    """The multi-rank receive path at kernel level: every fake rank reduces its own key splits with xattn_premerge into its block of
    a gathered buffer (the layout one all-gather of the reduced blocks produces), then xattn_merge2 reads the blocks in place with a
    NON-ZERO rank stride.  One rank holds no valid key for stream 1 (all LSE = -inf); one stream has a zero gate."""
    from vidi_b200 import ops
    rows = 7 * 4
    splits = [(5, 3), (4, 2), (6, 1), (2, 2), (1, 1), (3, 3), (2, 1), (4, 4)][:world]
    block = 2 * rows * (dh + 1)
    gathered = torch.full((world * block + 64,), float("nan"), device="cuda")
    per_stream = [[], []]
    for r, sp in enumerate(splits):
        srcs = []
        for si, P in enumerate(sp):
            O = rnd(P, rows, dh, seed=100 + 10 * r + si).float().contiguous()
            Ls = (3.0 * rnd(P, rows, seed=200 + 10 * r + si)).float().contiguous()
            Ls[0, ::5] = float("-inf")
            if r == 1 and si == 1:
                Ls[:] = float("-inf")
            srcs.append((O, Ls, P))
            per_stream[si].append(_merge_ref([(O[p].cpu(), Ls[p].cpu()) for p in range(P)]))
        pre = gathered[r * block:(r + 1) * block]
        ops.xattn_premerge(srcs, rows, dh, pre)
        for si in range(2):                      # the reduced partial itself
            o = pre[si * rows * (dh + 1):][:rows * dh].view(rows, dh).cpu().double()
            l = pre[si * rows * (dh + 1) + rows * dh:][:rows].cpu().double()
            Or, Lr = per_stream[si][-1]
            assert torch.allclose(o, Or, atol=2e-5, rtol=1e-4)
            fin = ~torch.isinf(Lr)
            assert torch.equal(torch.isinf(l), torch.isinf(Lr)) and torch.allclose(l[fin], Lr[fin], atol=2e-5, rtol=1e-5)
    att = rnd(rows, dh, seed=7).float().contiguous()
    gates = (1.0, 0.5)
    srcs = []
    for si in range(2):
        o = gathered[si * rows * (dh + 1):]
        srcs.append((o, o[rows * dh:], world, 1, block, block, gates[si]))
    out = torch.empty(rows, dh, device="cuda", dtype=BF)
    ops.xattn_merge2(srcs, att, out, rows, dh)
    ref = att.cpu().double() + sum(gates[si] * _merge_ref(per_stream[si])[0] for si in range(2))
    assert rel_err(out.cpu().double(), ref) < 4e-3
    assert not torch.isnan(gathered[:world * block]).any()          # every block fully written


def test_peer_exchange_single_process_group():
    """PartialExchange.local_group: `world` arenas on one device; every fake rank pushes its reduced partials into every arena and
    publishes its sequence number; merge (with the flag wait) on every rank gives the same bits and matches the fp64 reference.
    Run for several consecutive exchanges so that both slots and the sequence comparison are exercised."""
    from vidi_b200 import ops
    from vidi_b200.exchange import PartialExchange
    world, rows, dh = 4, 40, 256
    xs = PartialExchange.local_group(world, PartialExchange.capacity(rows, dh))
    for step in range(5):
        per_stream = [[], []]
        all_srcs = []
        for r in range(world):
            srcs = []
            for si, P in enumerate((3 + r, 1 + (r + step) % 3)):
                O = rnd(P, rows, dh, seed=1000 * step + 10 * r + si).float().contiguous()
                Ls = (2.0 * rnd(P, rows, seed=5000 + 1000 * step + 10 * r + si)).float().contiguous()
                if r == 2 and si == 0:
                    Ls[:] = float("-inf")
                srcs.append((O, Ls, P))
                per_stream[si].append(_merge_ref([(O[p].cpu(), Ls[p].cpu()) for p in range(P)]))
            all_srcs.append(srcs)
        for r in range(world):
            assert ops.xchg_push(xs[r], all_srcs[r], rows, dh) == step + 1
        att = rnd(rows, dh, seed=step).float().contiguous()
        outs = []
        for r in range(world):
            out = torch.empty(rows, dh, device="cuda", dtype=BF)
            ops.xchg_merge(xs[r], (1.0, 1.0), att, out, rows, dh)
            outs.append(out)
        torch.cuda.synchronize()
        for x in xs:
            x.check()
        ref = att.cpu().double() + sum(_merge_ref(per_stream[si])[0] for si in range(2))
        assert rel_err(outs[0].cpu().double(), ref) < 4e-3
        assert all(torch.equal(outs[0], o) for o in outs[1:])


def test_peer_exchange_missing_peer_sets_error_flag():
    """A rank that never publishes must surface as an error (bounded spin), not as a hung GPU."""
    from vidi_b200 import ops
    from vidi_b200.exchange import PartialExchange
    world, rows, dh = 2, 8, 128
    xs = PartialExchange.local_group(world, PartialExchange.capacity(rows, dh))
    O = rnd(2, rows, dh, seed=1).float().contiguous(); Ls = rnd(2, rows, seed=2).float().contiguous()
    ops.xchg_push(xs[0], [(O, Ls, 2)], rows, dh)          # rank 1 never pushes
    att = rnd(rows, dh, seed=3).float().contiguous()
    out = torch.empty(rows, dh, device="cuda", dtype=BF)
    ops.xchg_merge(xs[0], (1.0,), att, out, rows, dh)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="never delivered"):
        xs[0].check()


@pytest.mark.parametrize("dh,cap,G", [(256, 50.0, 2), (128, 0.0, 4)])
def test_xattn_splitkv_seg_equals_per_segment_launches(dh, cap, G):
    """One call / one launch over both key segments of a K||V cache (image rows | audio rows, the second with a key mask and a ragged
    end) must reproduce the per-segment launches bit for bit: same split sizes, same tiles, only the grid is shared.  dh=256 + soft-cap
    takes the tcgen05 kernel, dh=128 un-capped (Vidi-7B) the warp-level one."""
    from vidi_b200 import ops
    T, Hkv = 9, 2
    Hq = Hkv * G
    n0, n1 = 1500, 700
    kv = rnd(n0 + n1 + 37, 2 * Hkv * dh, seed=5).to(BF)             # rows after the segments must be ignored
    q = rnd(T, Hq * dh, seed=6).to(BF)
    mask1 = (torch.rand(n1, device="cuda") > 0.3).to(torch.uint8)
    kd = Hkv * dh
    splits = [3, 2]
    rows = T * Hq
    O = torch.full((sum(splits) * rows * dh,), float("nan"), device="cuda"); Ls = torch.full((sum(splits) * rows,), float("nan"), device="cuda")
    ops.xattn_splitkv_seg(q, kv[:, :kd], kv[:, kd:], [(0, n0, None), (n0, n1, mask1)], splits, Hq, Hkv, dh, dh ** -0.5, cap, O, Ls)
    o0, l0 = ops.xattn_splitkv(q, kv[:n0, :kd], kv[:n0, kd:], None, Hq, Hkv, dh, dh ** -0.5, cap, splits[0])
    o1, l1 = ops.xattn_splitkv(q, kv[n0:n0 + n1, :kd], kv[n0:n0 + n1, kd:], mask1, Hq, Hkv, dh, dh ** -0.5, cap, splits[1])
    assert torch.equal(O[:splits[0] * rows * dh], o0.reshape(-1)) and torch.equal(O[splits[0] * rows * dh:], o1.reshape(-1))
    assert torch.equal(Ls[:splits[0] * rows], l0.reshape(-1)) and torch.equal(Ls[splits[0] * rows:], l1.reshape(-1))
    # and the merged result against fp32 softmax over segment 1 alone (mask applied)
    out = torch.zeros(rows, dh, device="cuda")
    ops.xattn_merge(O[splits[0] * rows * dh:].view(splits[1], rows, dh), Ls[splits[0] * rows:].view(splits[1], rows), out)
    qf = q.float().view(T, Hq, dh); kf = kv[n0:n0 + n1, :kd].float().view(n1, Hkv, dh).repeat_interleave(G, 1)
    vf = kv[n0:n0 + n1, kd:].float().view(n1, Hkv, dh).repeat_interleave(G, 1)
    s_ = torch.einsum("thd,nhd->thn", qf, kf) * dh ** -0.5
    if cap > 0:
        s_ = cap * torch.tanh(s_ / cap)
    s_ = s_.masked_fill(mask1[None, None, :] == 0, float("-inf"))
    ref = torch.einsum("thn,nhd->thd", torch.softmax(s_, -1), vf).reshape(rows, dh)
    assert rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("case", ["late_spike", "masked_first_tile_then_very_negative", "benign"])
def test_xattn_uncapped_dh128_tcgen05_reference_window(case):
    """Vidi-7B cross attention (dh = 128, no soft-cap, Vidi_7B xattn.py:99-175) on the tcgen05 kernel: the per-row softmax reference is
    fixed after the first key tile.  'late_spike' puts logits ~180 nats above the first tile's max in a later tile and
    'masked_first_tile_then_very_negative' leaves the first tile without a valid key and all later logits ~ -150: both must take the
    in-kernel second pass and still match fp32 softmax and the warp-level kernel."""
    from vidi_b200 import ops
    T, Hkv, G, dh, N = 5, 2, 4, 128, 640
    Hq = Hkv * G
    q = rnd(T, Hq * dh, seed=11).to(BF)
    kv = rnd(N, 2 * Hkv * dh, seed=12).to(BF)
    kd = Hkv * dh
    mask = torch.ones(N, device="cuda", dtype=torch.uint8)
    scale = dh ** -0.5
    if case == "late_spike":
        # keys 300..303 are aligned with the queries of head 0 / token 0 and scaled up: logits ~ +180
        kv[300:304, :dh] = (q[0, :dh].float() * 14.0).to(BF)
    elif case == "masked_first_tile_then_very_negative":
        mask[:64] = 0
        kv[:, :kd] = (-q[0, :dh].float().repeat(Hkv) * 12.0).to(BF)[None, :]        # every key anti-aligned with token 0 / head 0
    out = {}
    for impl in ("auto", "mma"):
        o, l = ops.xattn_splitkv(q, kv[:, :kd], kv[:, kd:], mask, Hq, Hkv, dh, scale, 0.0, 2, impl=impl)
        m = torch.zeros(T * Hq, dh, device="cuda")
        ops.xattn_merge(o, l, m)
        out[impl] = m
    qf = q.float().view(T, Hq, dh); kf = kv[:, :kd].float().view(N, Hkv, dh).repeat_interleave(G, 1)
    vf = kv[:, kd:].float().view(N, Hkv, dh).repeat_interleave(G, 1)
    s_ = (torch.einsum("thd,nhd->thn", qf, kf) * scale).masked_fill(mask[None, None, :] == 0, float("-inf"))
    ref = torch.einsum("thn,nhd->thd", torch.softmax(s_, -1), vf).reshape(T * Hq, dh)
    if case != "benign":
        assert float(s_[torch.isfinite(s_)].abs().max()) > 100.0          # the case really leaves the +-100 (log2) window
    assert torch.isfinite(out["auto"]).all()
    assert rel_err(out["auto"], ref) < 1e-2, rel_err(out["auto"], ref)
    assert rel_err(out["mma"], ref) < 1e-2


@pytest.mark.parametrize("M,N,K,glu,act", [
    (1, 8192, 3584, 0, 0), (32, 8192, 3584, 0, 0), (32, 3584, 4096, 0, 0), (7, 3584, 14336, 0, 0), (24, 28672, 3584, 1, 0),
    (1, 28672, 3584, 1, 0), (40, 1024, 512, 0, 0), (64, 2048, 512, 2, 0), (5, 25600, 3584, 0, 3), (33, 520, 288, 0, 0), (16, 4096, 1032, 0, 0)])
def test_gemm_skinny_text_shapes(M, N, K, glu, act):
    """Weight-streaming GEMM of the text stream (gemm_skinny_sm100.cu: swapped operands, split-K with fixed-order reduction): against
    fp32 torch and against the general kernel (block_n < 0 forces it); two launches must be bit-identical (deterministic reduction)."""
    from vidi_b200 import ops
    from vidi_b200.weights import pack_glu
    a = rnd(M, K, seed=71).to(BF)
    if glu:
        wg = rnd(N // 2, K, scale=0.03, seed=72).to(BF); wu = rnd(N // 2, K, scale=0.03, seed=73).to(BF)
        w = pack_glu(wg, wu, 256)
        g = a.float() @ wg.float().t()
        ref = (F.gelu(g, approximate="tanh") if glu == 1 else F.silu(g)) * (a.float() @ wu.float().t())
    else:
        w = rnd(N, K, scale=0.03, seed=72).to(BF)
        ref = a.float() @ w.float().t()
        if act == 3:
            ref = 30.0 * torch.tanh(ref / 30.0)
    kw = dict(glu=glu, act=act, act_param=30.0, out_fp32=(act == 3))
    out = ops.gemm(a, w, **kw)
    out2 = ops.gemm(a, w, **kw)
    gen = ops.gemm(a, w, block_n=-(256 if glu else 64), **kw)
    torch.cuda.synchronize()
    assert out.shape == ref.shape and torch.equal(out, out2)
    assert rel_err(out, ref) < 6e-3, rel_err(out, ref)
    assert rel_err(out, gen) < 6e-3, rel_err(out, gen)
    assert float((out.float() - ref).abs().max()) < 0.05 * float(ref.abs().max()) + 1e-2

"""Device-side frame pre-processing (vidi_resample_u8 / vidi_resample_u8_to_chw_bf16) against the torch restatement that
tests/test_preprocess_cpu.py pins bit-exactly to Pillow.

Round 2: first run on a B200 showed the integer passes exact and the bf16 output off by one ulp in places (the affine had been
contracted into an FMA); fixed with explicitly rounded mul / sub / div, and the tests are unconditional."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F,H,W,S", [(3, 270, 480, 384), (2, 720, 1280, 384), (2, 384, 384, 384), (1, 100, 60, 384), (2, 500, 384, 384),
                                     (1, 384, 777, 384), (1, 1080, 1920, 384), (4, 90, 160, 48)])
def test_resize_frames_bit_exact(F, H, W, S):
    from vidi_b200 import ops
    from vidi_b200.preprocess import SiglipImageProcessorLite, resize_bicubic_u8
    g = torch.Generator().manual_seed(H * 3 + W)
    frames = torch.randint(0, 256, (F, H, W, 3), generator=g, dtype=torch.uint8)
    frames[:, : H // 3] = (frames[:, : H // 3] // 64) * 64
    fc = frames.cuda()
    # the uint8 intermediate of the horizontal pass alone
    if W != S:
        from vidi_b200 import lib as _lib
        xmin, kk = ops._resample_tables(W, S, torch.device("cuda"))
        mid = torch.empty(F, H, S, 3, device="cuda", dtype=torch.uint8)
        L = _lib.load()
        rc = L.vidi_resample_u8(ops._ptr(fc), ops._ptr(mid), F * H, W, S, 3, ops._ptr(xmin), ops._ptr(kk), kk.shape[1], ops._stream())
        assert rc == 0
        from vidi_b200.preprocess import _resample_axis_u8
        assert torch.equal(mid.cpu(), _resample_axis_u8(frames, S, 2))
    out = ops.resize_frames_u8(fc, S)
    ref = SiglipImageProcessorLite(S).preprocess(frames).to(torch.bfloat16)
    assert out.shape == ref.shape
    bad = int((out.cpu() != ref).sum())
    assert bad == 0, f"{bad} of {ref.numel()} bf16 outputs differ, max |d| = {float((out.cpu().float() - ref.float()).abs().max())}"



@pytest.mark.parametrize("seconds", [7.3, 61.7])
def test_log_mel_device_matches_torch_restatement(seconds):
    """ops.log_mel (framing / power / finish kernels around two split-bf16 tensor-core GEMMs) vs WhisperFeatureExtractorLite.log_mel
    (pinned to the HF extractor at 1e-4 by tests/test_preprocess_cpu.py); tolerance 1e-3 + bf16 output rounding."""
    import math
    from vidi_b200 import ops
    from vidi_b200.preprocess import WhisperFeatureExtractorLite
    n = int(seconds * 16000)
    t = torch.arange(n) / 16000.0
    g = torch.Generator().manual_seed(n)
    audio = 0.3 * torch.sin(2 * math.pi * 440 * t) * ((t % 3) < 2) + 0.05 * torch.randn(n, generator=g)
    fe = WhisperFeatureExtractorLite(128)
    C = -(-n // fe.n_samples)
    buf = torch.zeros(C * fe.n_samples)
    buf[:n] = audio
    ref = fe.log_mel(buf.view(C, -1))
    out = ops.log_mel(buf.view(C, -1).cuda())
    assert out.shape == (C, 128, 3000) and out.dtype == torch.bfloat16
    err = (out.float().cpu() - ref).abs()
    assert float(err.max()) <= 1e-3 + 2 ** -8 * float(ref.abs().max()), float(err.max())


@pytest.mark.parametrize("poly", [2, 3, 4])
@pytest.mark.parametrize("B,S,H,dh", [(3, 729, 4, 72), (2, 1500, 4, 64), (2, 300, 2, 64)])
def test_attn_dense_poly_exp2_variants(B, S, H, dh, poly):
    """parked A/B variants of the tower attention with an FMA-pipe exp2 polynomial on every poly-th score pair (ops.attn_dense impl="polyN")"""
    from vidi_b200 import ops
    d = H * dh
    g = torch.Generator(device="cuda").manual_seed(S + poly)
    qkv = torch.randn(B * S, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
    out = ops.attn_dense(qkv, B, S, H, dh, dh ** -0.5, impl=f"poly{poly}")
    q, k, v = [t.float().view(B, S, H, dh).transpose(1, 2) for t in qkv.split(d, dim=1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v).transpose(1, 2).reshape(B * S, d)
    err = float((out.float() - ref).norm() / ref.norm())
    assert err < 8e-3, err

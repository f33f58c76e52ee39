"""Host / torch-op side of the reference's ``ask()`` pre-processing (8f "next"):

  * frames: ``process_images`` 'resize' branch (Vidi1.5_9B/vidi/dataset/img_utils.py:181-187) =
    ``PIL.Image.resize((S, S), BICUBIC)`` on uint8 RGB, then the HF SiglipImageProcessor affine (x/255 - 0.5)/0.5.
    ``resize_bicubic_u8`` restates Pillow's resampler (libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
    ImagingResampleHorizontal/Vertical_8bpc) in integer arithmetic — 22-bit fixed-point coefficients, a uint8 intermediate
    between the horizontal and the vertical pass, floor-shift + clamp — and is BIT-EXACT against Pillow
    (tests/test_preprocess_cpu.py).  It is written with torch ops only, so the same code runs on a CUDA tensor.
  * audio: ``process_audio`` (vid_utils.py:52-63) = 30-s chunks -> HF WhisperFeatureExtractor log-mel ([C,128,3000]) and
    ``audio_size = sum floor(len_chunk / 160)``.  ``log_mel`` restates that extractor (hann window 400, hop 160, reflect-padded
    centred STFT, power spectrum, Slaney mel filter bank, log10 clamp 1e-10, last frame dropped, max-8 dB floor PER CHUNK,
    (x + 4) / 4) with torch.stft; tolerance vs the HF numpy path 1e-4 absolute (the HF docstring itself quotes 1e-5
    between its numpy and torch paths).

Video / audio DECODING (decord, ffmpeg: vid_utils.py:9-49) is out of scope — the boundary here starts at decoded uint8 frames
and 16 kHz float32 mono samples.  These run before the timed prefill path (bench inputs are already pre-processed tensors).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch

_PRECISION_BITS = 32 - 8 - 2            # Resample.c: coefficients are 22-bit fixed point for 8-bit channels


def _bicubic(x: float) -> float:
    """Resample.c bicubic_filter with a = -0.5 (Keys / Catmull-Rom)."""
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def pil_bicubic_coeffs(in_size: int, out_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole axis (box = [0, in_size]).
    Returns (xmin int64 [out], kk int64 [out, ksize]); taps beyond a row's xmax are zero."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale                              # bicubic support = 2
    ksize = int(math.ceil(support)) * 2 + 1
    xmins, rows = [], []
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)                   # C (int) cast: truncation toward zero
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        k = [int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS)) for v in w]
        rows.append(k + [0] * (ksize - xmax))
        xmins.append(xmin)
    return torch.tensor(xmins, dtype=torch.int64), torch.tensor(rows, dtype=torch.int64)


def _resample_axis_u8(x: torch.Tensor, out_size: int, axis: int) -> torch.Tensor:
    """one 8-bit resampling pass along ``axis`` of a uint8 tensor (ImagingResampleHorizontal_8bpc / Vertical_8bpc)."""
    in_size = x.shape[axis]
    xmin, kk = pil_bicubic_coeffs(in_size, out_size)
    xmin, kk = xmin.to(x.device), kk.to(x.device)
    ksize = kk.shape[1]
    idx = (xmin[:, None] + torch.arange(ksize, device=x.device)[None, :]).clamp_(max=in_size - 1)   # zero taps where clamped
    xm = x.movedim(axis, -1)                                                                       # [..., in]
    acc = torch.full(xm.shape[:-1] + (out_size,), 1 << (_PRECISION_BITS - 1), dtype=torch.int64, device=x.device)
    for k in range(ksize):                                    # ksize is 5 (upscale) .. ~25 (5x downscale): a short loop
        acc += xm[..., idx[:, k]].to(torch.int64) * kk[:, k]
    out = (acc >> _PRECISION_BITS).clamp_(0, 255).to(torch.uint8)                                    # clip8: floor shift, clamp
    return out.movedim(-1, axis)


def resize_bicubic_u8(frames: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """frames uint8 [..., H, W, C] -> uint8 [..., out_h, out_w, C], bit-exact ``PIL.Image.resize((out_w, out_h), BICUBIC)``:
    horizontal pass first, rounded to uint8, then the vertical pass; a pass whose size does not change is skipped."""
    assert frames.dtype == torch.uint8 and frames.dim() >= 3
    x = frames
    if x.shape[-2] != out_w:
        x = _resample_axis_u8(x, out_w, x.dim() - 2)
    if x.shape[-3] != out_h:
        x = _resample_axis_u8(x, out_h, x.dim() - 3)
    return x


class SiglipImageProcessorLite:
    """The slice of HF SiglipImageProcessor the reference touches: ``size``, ``image_mean`` / ``image_std`` (0.5), ``output_size``
    (img_utils.py:184) and ``preprocess`` = resize -> x/255 -> (x - 0.5) / 0.5, output [F,3,S,S] fp32."""
    image_mean = [0.5, 0.5, 0.5]
    image_std = [0.5, 0.5, 0.5]
    rescale_factor = 1.0 / 255.0

    def __init__(self, size: int):
        self.size = {"height": size, "width": size}
        self.output_size = size

    def preprocess(self, frames_uint8, return_tensors=None, chunk: int = 64):
        """frames [F,H,W,3] uint8 (decoded RGB) -> [F,3,S,S] float32.
        Called the reference's way — ``preprocess(pil_image, return_tensors='pt')['pixel_values'][0]`` (img_utils.py:180,186) — it
        accepts a PIL image (or a list of them) and returns ``{"pixel_values": [n,3,S,S]}``."""
        if not torch.is_tensor(frames_uint8):
            import numpy as np
            imgs = frames_uint8 if isinstance(frames_uint8, (list, tuple)) else [frames_uint8]
            out = [self.preprocess(torch.from_numpy(np.asarray(im.convert("RGB"), dtype=np.uint8).copy())[None]) for im in imgs]
            return {"pixel_values": torch.cat(out, 0)}
        assert frames_uint8.dtype == torch.uint8 and frames_uint8.dim() == 4 and frames_uint8.shape[-1] == 3
        s = self.output_size
        if frames_uint8.is_cuda:
            # device path (csrc/preproc.cu): Pillow's two 8-bit passes as kernels, the second fused with the affine and the CHW layout
            # change; output is bf16 (what the engine consumes; bit-equal to this method's CPU result rounded to bf16)
            from . import ops
            return ops.resize_frames_u8(frames_uint8.contiguous(), s, self.rescale_factor, self.image_mean[0], self.image_std[0])
        outs = []
        for i in range(0, frames_uint8.shape[0], chunk):
            r = resize_bicubic_u8(frames_uint8[i:i + chunk], s, s)
            outs.append(((r.permute(0, 3, 1, 2).to(torch.float32) * self.rescale_factor) - 0.5) / 0.5)
        return torch.cat(outs, 0) if outs else torch.empty(0, 3, s, s)


# ----------------------------------------------------------------------------------------------------------------
# Whisper log-mel (HF WhisperFeatureExtractor: feature_size 128, n_fft 400, hop 160, 30-s chunks, 16 kHz)
# ----------------------------------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f: torch.Tensor) -> torch.Tensor:
    min_log_hz, min_log_mel, logstep = 1000.0, 15.0, 27.0 / math.log(6.4)
    mel = 3.0 * f / 200.0
    return torch.where(f >= min_log_hz, min_log_mel + torch.log(f.clamp(min=1e-10) / min_log_hz) * logstep, mel)


def _mel_to_hz_slaney(m: torch.Tensor) -> torch.Tensor:
    min_log_hz, min_log_mel, logstep = 1000.0, 15.0, math.log(6.4) / 27.0
    return torch.where(m >= min_log_mel, min_log_hz * torch.exp(logstep * (m - min_log_mel)), 200.0 * m / 3.0)


def mel_filter_bank(n_freqs: int = 201, n_mels: int = 128, sr: int = 16000) -> torch.Tensor:
    """transformers.audio_utils.mel_filter_bank(num_frequency_bins, num_mel_filters, 0, sr/2, sr, norm="slaney",
    mel_scale="slaney") -> float64 [n_freqs, n_mels] triangular filters with Slaney area normalisation."""
    fft_freqs = torch.linspace(0, sr // 2, n_freqs, dtype=torch.float64)
    m_lo, m_hi = _hz_to_mel_slaney(torch.tensor(0.0, dtype=torch.float64)), _hz_to_mel_slaney(torch.tensor(sr / 2.0, dtype=torch.float64))
    filt_freqs = _mel_to_hz_slaney(torch.linspace(float(m_lo), float(m_hi), n_mels + 2, dtype=torch.float64))
    fdiff = filt_freqs[1:] - filt_freqs[:-1]
    slopes = filt_freqs[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = torch.clamp(torch.minimum(down, up), min=0.0)
    enorm = 2.0 / (filt_freqs[2:n_mels + 2] - filt_freqs[:n_mels])
    return fb * enorm[None, :]


class WhisperFeatureExtractorLite:
    """The slice of HF WhisperFeatureExtractor that ``process_audio`` (vid_utils.py:52-63) uses."""
    sampling_rate = 16000
    n_fft = 400
    hop_length = 160
    chunk_length = 30
    n_samples = 480000
    nb_max_frames = 3000

    def __init__(self, mels: int = 128):
        self.feature_size = mels
        self.mel_filters = mel_filter_bank(1 + self.n_fft // 2, mels, self.sampling_rate)

    def audio_size(self, n_samples: int) -> int:
        """audio_size = sum over 30-s chunks of floor(len/160) (vid_utils.py:62 ``num_frames.sum()``)."""
        full, rest = divmod(n_samples, self.n_samples)
        return full * (self.n_samples // self.hop_length) + rest // self.hop_length

    def log_mel(self, chunks: torch.Tensor) -> torch.Tensor:
        """chunks [C, 480000] float32 (zero-padded 30-s windows) -> [C, mels, 3000] float32."""
        dev = chunks.device
        window = torch.hann_window(self.n_fft, periodic=True, dtype=torch.float64, device=dev)
        st = torch.stft(chunks.to(torch.float64), self.n_fft, self.hop_length, window=window, center=True, pad_mode="reflect",
                        return_complex=True)
        power = st.real ** 2 + st.imag ** 2                                        # [C, 201, 3001]
        mel = self.mel_filters.to(dev).t() @ power                                  # [C, mels, 3001]
        log_spec = torch.log10(torch.clamp(mel, min=1e-10))[..., :-1]
        log_spec = torch.maximum(log_spec, log_spec.amax(dim=(1, 2), keepdim=True) - 8.0)
        return ((log_spec + 4.0) / 4.0).to(torch.float32)

    def __call__(self, audio, sampling_rate=None, return_tensors=None, return_token_timestamps=False, **kw):
        """audio [n] float32 mono 16 kHz -> (input_features [C,mels,3000], audio_size) exactly as ``process_audio`` returns them.
        Called the reference's way — a LIST of <= 30-s numpy chunks with ``sampling_rate=, return_tensors='pt',
        return_token_timestamps=True`` (vid_utils.py:57-61) — it returns an object with ``.input_features`` and ``.num_frames``."""
        if isinstance(audio, (list, tuple)):
            from types import SimpleNamespace
            assert sampling_rate in (None, self.sampling_rate), "WhisperFeatureExtractorLite: 16 kHz input only"
            buf = torch.zeros(len(audio), self.n_samples, dtype=torch.float32)
            for i, c in enumerate(audio):
                c = torch.as_tensor(c, dtype=torch.float32)
                assert c.dim() == 1 and c.numel() <= self.n_samples
                buf[i, :c.numel()] = c
            frames = torch.tensor([len(c) // self.hop_length for c in audio], dtype=torch.int64)
            return SimpleNamespace(input_features=self.log_mel(buf), num_frames=frames)
        assert audio.dim() == 1
        n = audio.numel()
        C = max(1, -(-n // self.n_samples))
        buf = torch.zeros(C * self.n_samples, dtype=torch.float32, device=audio.device)
        buf[:n] = audio.to(torch.float32)
        if audio.is_cuda:
            # device path (csrc/preproc.cu + two split-bf16 tensor-core GEMMs): bf16 features, <= 1e-3 from the fp64 STFT below
            from . import ops
            return ops.log_mel(buf.view(C, self.n_samples), self.feature_size), self.audio_size(n)
        return self.log_mel(buf.view(C, self.n_samples)), self.audio_size(n)


def sample_frame_indices(n_frames: int, avg_fps: float, fps: float = 1.0) -> List[int]:
    """``load_video`` without a time range (vid_utils.py:11-13): every round(avg_fps / fps)-th frame."""
    step = round(avg_fps / fps)
    return list(range(0, n_frames, step))

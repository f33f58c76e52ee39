"""Minimal host-side stand-ins for the HF processors the reference's ``ask()`` touches
(vidi/dataset/img_utils.py:173-198, vid_utils.py:52-63).  Only the attributes / tensor contracts the
prefill path needs; full device-side preprocessing is a "next" row of SURVEY.md 8(f)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


class SiglipImageProcessorLite:
    """resize -> [0,1] -> (x-0.5)/0.5, output [F,3,S,S] fp32 (img_utils.py:181-187 'resize' branch)."""
    image_mean = [0.5, 0.5, 0.5]
    image_std = [0.5, 0.5, 0.5]

    def __init__(self, size: int):
        self.size = {"height": size, "width": size}
        self.output_size = size

    def preprocess(self, frames_uint8: torch.Tensor) -> torch.Tensor:
        """frames [F,H,W,3] uint8 -> [F,3,S,S] float."""
        x = frames_uint8.permute(0, 3, 1, 2).float() / 255.0
        s = self.output_size
        x = F.interpolate(x, size=(s, s), mode="bicubic", align_corners=False).clamp_(0, 1)
        return (x - 0.5) / 0.5


class WhisperFeatureExtractorLite:
    sampling_rate = 16000
    nb_max_frames = 3000
    hop_length = 160
    chunk_length = 30

    def __init__(self, mels: int = 128):
        self.feature_size = mels

    def audio_size(self, n_samples: int) -> int:
        """audio_size = sum floor(len/160) (vid_utils.py:62; SURVEY.md 8c note on transformers 5.x)."""
        return n_samples // self.hop_length

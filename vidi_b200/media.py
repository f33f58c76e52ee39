"""Media decoding in front of the prefill path: ``load_video`` / ``load_audio`` / ``get_length`` of the reference
(Vidi1.5_9B/vidi/dataset/vid_utils.py:9-49, vidi/eval/inference.py:68-75) without decord: frames come from OpenCV's FFmpeg backend,
audio from the ``ffmpeg`` CLI exactly as the reference invokes it when the binary exists, or from a RIFF/WAVE file through the standard
library.  Decoding is host work and NOT part of the timed hot path; the decoded uint8 frames / float32 samples can be moved to the GPU
and handed to ``pipeline.ask`` (device-side resize + log-mel).  ``ask_path`` is the reference's ``ask(question, vid_path, ...)``."""
from __future__ import annotations

import shutil
import subprocess
import wave
from typing import Optional, Tuple

import numpy as np
import torch


def load_video(file: str, fps: float = 1.0, time_range: Optional[Tuple[float, float]] = None) -> torch.Tensor:
    """-> uint8 RGB frames [F, H, W, 3] sampled as vid_utils.py:9-22 does: every round(avg_fps / fps)-th frame from frame 0, or
    ``round((t1 - t0) * fps)`` frames linearly spaced over the frame-index range of ``time_range``."""
    import cv2
    cap = cv2.VideoCapture(str(file))
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open video {file}")
    n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    avg_fps = float(cap.get(cv2.CAP_PROP_FPS))
    if time_range is None:
        step = max(1, round(avg_fps / fps))
        wanted = list(range(0, n, step))
    else:
        i0 = round(time_range[0] * avg_fps)
        i1 = min(round(time_range[1] * avg_fps), n - 1)
        wanted = np.linspace(i0, i1, round((time_range[1] - time_range[0]) * fps), dtype=int).tolist()
    frames, want, idx = [], set(wanted), 0
    last = max(wanted) if wanted else -1
    while idx <= last:                                # sequential decode: exact frame indices for any codec / GOP structure
        ok, bgr = cap.read()
        if not ok:
            break
        if idx in want:
            frames.append(torch.from_numpy(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)))
        idx += 1
    cap.release()
    if not frames:
        raise ValueError(f"no frame decoded from {file}")
    by_idx = dict(zip(sorted(want), frames))          # linspace may repeat an index: repeat the frame like decord's get_batch
    return torch.stack([by_idx[i] for i in wanted if i in by_idx])


def load_audio(file: str, sample_rate: int = 16000, time_range: Optional[Tuple[float, float]] = None) -> torch.Tensor:
    """-> float32 mono samples at ``sample_rate`` (vid_utils.py:25-49: whisper's loading recipe through the ffmpeg CLI)."""
    if shutil.which("ffmpeg"):
        cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", str(file)]
        if time_range is not None:
            cmd += ["-ss", f"{time_range[0]:.2f}", "-t", f"{time_range[1] - time_range[0]:.2f}"]
        cmd += ["-f", "s16le", "-ac", "1", "-acodec", "pcm_s16le", "-ar", str(sample_rate), "-"]
        out = subprocess.run(cmd, capture_output=True, check=True).stdout
        return torch.from_numpy(np.frombuffer(out, np.int16).flatten().astype(np.float32) / 32768.0)
    if str(file).lower().endswith(".wav"):
        with wave.open(str(file), "rb") as w:
            if w.getsampwidth() != 2 or w.getframerate() != sample_rate:
                raise ValueError(f"{file}: need 16-bit PCM at {sample_rate} Hz when ffmpeg is not installed")
            pcm = np.frombuffer(w.readframes(w.getnframes()), np.int16).reshape(-1, w.getnchannels())
        mono = pcm.astype(np.float32).mean(1) / 32768.0
        if time_range is not None:
            mono = mono[int(time_range[0] * sample_rate):int(time_range[1] * sample_rate)]
        return torch.from_numpy(np.ascontiguousarray(mono))
    raise RuntimeError("audio decoding needs the ffmpeg binary (as the reference does, vid_utils.py:28-46) or a 16 kHz 16-bit .wav file")


def get_length(file: str) -> float:
    """media duration in seconds (inference.py:68-75 asks ffprobe; without it: frame count / fps from the container)"""
    if shutil.which("ffprobe"):
        r = subprocess.run(["ffprobe", "-v", "error", "-show_entries", "format=duration", "-of", "default=noprint_wrappers=1:nokey=1", str(file)],
                           capture_output=True, text=True)
        try:
            return float(r.stdout)
        except ValueError:
            pass
    import cv2
    cap = cv2.VideoCapture(str(file))
    n, f = cap.get(cv2.CAP_PROP_FRAME_COUNT), cap.get(cv2.CAP_PROP_FPS)
    cap.release()
    if f <= 0:
        raise ValueError(f"cannot read the duration of {file}")
    return float(n / f)


def ask_path(question: str, vid_path: str, model, tokenizer, image_processor, audio_processor, family: str = "vidi15",
             audio_path: Optional[str] = None, device: Optional[str] = "cuda", max_new_tokens: int = 1024) -> str:
    """``ask(question, vid_path, model, tokenizer, image_processor, audio_processor)`` of inference.py:18-66 from a file path: decode on
    the host, pre-process on the GPU (``device="cuda"``) or on the host (None), generate, format.  ``audio_path`` lets the sound come
    from a separate .wav when ffmpeg is not installed."""
    from .pipeline import ask
    frames = load_video(vid_path)
    audio = load_audio(audio_path or vid_path, audio_processor.sampling_rate)
    length = get_length(vid_path)
    if device is not None:
        frames, audio = frames.to(device), audio.to(device)
    return ask(question, frames, audio, length, model, tokenizer, image_processor, audio_processor, family=family, max_new_tokens=max_new_tokens)

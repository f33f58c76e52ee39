"""Frozen configuration for the Vidi prefill path.

Mirrors the keys the reference reads from the checkpoint's ``config.json``
(``DattnGemma2Config``, Vidi1.5_9B/vidi/model/lmm/dattn/gemma.py:427-448;
``mm_*`` overrides, Vidi1.5_9B/scripts/finetune.sh:18-25; ``mm_time_interval``
default 10000, Vidi1.5_9B/vidi/train/train.py:51) plus the public HF dims of the
three towers.  No env switches.
"""
from __future__ import annotations

import dataclasses
import json
import math
from dataclasses import dataclass, field


@dataclass(frozen=True)
class VisionCfg:
    """SigLIP-so400m/14@384 vision tower (HF SiglipVisionConfig)."""
    hidden: int = 1152
    heads: int = 16
    inter: int = 4304
    layers: int = 27          # checkpoint layers; hidden_states[select_layer=-2] => layers-1 are run
    image: int = 384
    patch: int = 14
    eps: float = 1e-6
    select_layer: int = -2

    @property
    def side(self) -> int:
        return self.image // self.patch

    @property
    def patches(self) -> int:
        return self.side * self.side

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def run_layers(self) -> int:
        # hidden_states has layers+1 entries (embeddings first); index -2 == output of layer `layers-1`
        return self.layers + 1 + self.select_layer


@dataclass(frozen=True)
class AudioCfg:
    """Whisper-large-v3 encoder (HF WhisperConfig, encoder half only)."""
    d_model: int = 1280
    heads: int = 20
    ffn: int = 5120
    layers: int = 32
    mels: int = 128
    max_source_positions: int = 1500
    nb_max_frames: int = 3000
    eps: float = 1e-5

    @property
    def head_dim(self) -> int:
        return self.d_model // self.heads


@dataclass(frozen=True)
class LLMCfg:
    """Gemma2-9B decoder dims (HF Gemma2Config)."""
    hidden: int = 3584
    heads: int = 16
    kv_heads: int = 8
    head_dim: int = 256
    inter: int = 14336
    layers: int = 42
    vocab: int = 256000
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    query_pre_attn_scalar: float = 256.0
    attn_softcap: float = 50.0
    final_softcap: float = 30.0
    sliding_window: int = 4096
    tie_word_embeddings: bool = True

    @property
    def q_dim(self) -> int:
        return self.heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.kv_heads * self.head_dim

    @property
    def groups(self) -> int:
        return self.heads // self.kv_heads


@dataclass(frozen=True)
class Vidi15Config:
    """Vidi1.5-9B: Gemma2 Dattn LMM + SigLIP + Whisper."""
    llm: LLMCfg = field(default_factory=LLMCfg)
    vis: VisionCfg = field(default_factory=VisionCfg)
    aud: AudioCfg = field(default_factory=AudioCfg)
    mm_image_pool_size: int = 2
    mm_audio_pool_size: int = 5
    mm_time_interval: int = 10000
    mm_std: float = 0.028976401314139366
    mm_eps: float = 1e-5                 # vidi/model/mm_layer/norm.py:9,19
    max_image_tokens: int = 60000        # multimodal.py:176
    mm_splits: int = 1                   # kept settable for drop-in (inference.py:86); ignored
    mm_image_aspect_ratio: str = "resize"   # read by process_images(video, image_processor, model.config) (img_utils.py:173-198)
    name: str = "vidi1.5-9b"

    # ---- token math (multimodal.py:175-180, utils.py:152-171) ----
    def image_hw(self, n_frames: int) -> tuple[int, int]:
        """Feature-map size after pad(27->28) and the optional bilinear shrink."""
        side = self.vis.side + 1
        n_tokens = n_frames * side * side
        max_tokens = self.max_image_tokens * self.mm_image_pool_size ** 2
        if n_tokens > max_tokens:
            ratio = math.sqrt(max_tokens / (n_frames * side * side))
            th, tw = int(side * ratio), int(side * ratio)
            return max(10, th - th % 2), max(10, tw - tw % 2)
        return 28, 28

    def image_tokens(self, n_frames: int) -> int:
        h, w = self.image_hw(n_frames)
        p = self.mm_image_pool_size
        return n_frames * (h // p) * (w // p)

    def audio_tokens(self, audio_size: int) -> int:
        r = self.aud.max_source_positions / self.aud.nb_max_frames
        return int(math.floor(math.floor(audio_size * r) / self.mm_audio_pool_size))

    def to_json(self) -> str:
        return json.dumps(dataclasses.asdict(self), indent=1)


def vidi15_9b() -> Vidi15Config:
    return Vidi15Config()


def vidi15_mini(llm_layers: int = 2, vis_layers: int = 3, aud_layers: int = 2) -> Vidi15Config:
    """Small dims that keep the true head dims (256 / 72 / 64) so the sm_100a kernels
    are exercised with their production template parameters."""
    return Vidi15Config(
        llm=LLMCfg(hidden=512, heads=4, kv_heads=2, head_dim=256, inter=1024, layers=llm_layers, vocab=1024),
        vis=VisionCfg(hidden=288, heads=4, inter=520, layers=vis_layers, image=378, patch=14),
        aud=AudioCfg(d_model=256, heads=4, ffn=512, layers=aud_layers),
        name="vidi1.5-mini",
    )


def vidi15_true_dims(llm_layers: int = 2, vis_layers: int = 3, aud_layers: int = 2, vocab: int = 8192) -> Vidi15Config:
    """True hidden dims of the 9B model with the depth cut, for GPU parity against the CPU oracle."""
    return Vidi15Config(
        llm=dataclasses.replace(LLMCfg(), layers=llm_layers, vocab=vocab),
        vis=dataclasses.replace(VisionCfg(), layers=vis_layers),
        aud=dataclasses.replace(AudioCfg(), layers=aud_layers),
        name=f"vidi1.5-9b-dims-L{llm_layers}",
    )


# ---------------------------------------------------------------------------------------
# Vidi-7B (Mistral) -- Vidi_7B/model/lmm/dattn/mistral.py:456-477
# ---------------------------------------------------------------------------------------
@dataclass(frozen=True)
class MistralCfg:
    hidden: int = 4096
    heads: int = 32
    kv_heads: int = 8
    head_dim: int = 128
    inter: int = 14336
    layers: int = 32
    vocab: int = 32000
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    tie_word_embeddings: bool = False
    sliding_window: int = 0          # MistralConfig.sliding_window (4096 in Mistral-7B-v0.1, null in v0.2+); 0 = full attention

    @property
    def q_dim(self) -> int:
        return self.heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.kv_heads * self.head_dim

    @property
    def groups(self) -> int:
        return self.heads // self.kv_heads


@dataclass(frozen=True)
class Vidi7BConfig:
    llm: MistralCfg = field(default_factory=MistralCfg)
    vis: VisionCfg = field(default_factory=VisionCfg)
    aud: AudioCfg = field(default_factory=AudioCfg)
    mm_image_pool_size: int = 16     # tokens/frame = pool^2; read from ckpt config.json when present
    mm_audio_pool_size: int = 5
    mm_time_interval: int = 10000
    mm_std: float = 0.028976401314139366
    mm_eps: float = 1e-5
    mm_splits: int = 1
    mm_image_aspect_ratio: str = "resize"
    name: str = "vidi-7b"

    def image_tokens(self, n_frames: int) -> int:
        return n_frames * self.mm_image_pool_size ** 2

    def audio_tokens(self, audio_size: int) -> int:
        r = self.aud.max_source_positions / self.aud.nb_max_frames
        return int(math.floor(math.floor(audio_size * r) / self.mm_audio_pool_size))


def vidi7b_mini(llm_layers: int = 2, vis_layers: int = 3, aud_layers: int = 2) -> Vidi7BConfig:
    return Vidi7BConfig(
        llm=MistralCfg(hidden=512, heads=8, kv_heads=2, head_dim=128, inter=1024, layers=llm_layers, vocab=1024),
        vis=VisionCfg(hidden=288, heads=4, inter=520, layers=vis_layers, image=378, patch=14),
        aud=AudioCfg(d_model=256, heads=4, ffn=512, layers=aud_layers),
        mm_image_pool_size=4,
        name="vidi-7b-mini",
    )

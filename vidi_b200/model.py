"""Drop-in Python surface of the reference's L3/L4 layer for the inference prefill path.

Mirrors (same names, argument meaning, error behaviour):
  * ``load_pretrained_model``                      Vidi1.5_9B/vidi/model/builder.py:24-64
  * ``DattnGemma2ForCausalLM.forward / generate``  Vidi1.5_9B/vidi/model/lmm/dattn/gemma.py:484-601, 603-655
  * ``DattnCausalLMOutputWithPast``                Vidi1.5_9B/vidi/model/lmm/dattn/outputs.py:11-19
  * ``IMAGE_TOKEN_INDEX`` handling                 Vidi1.5_9B/vidi/model/lmm/dattn/multimodal.py:339-451
so ``vidi/eval/inference.py::ask`` runs unchanged on top of it.  Everything below
the facade is the sm_100a engine; there is no CPU / eager fallback.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional

import torch

from .config import Vidi15Config, Vidi7BConfig, LLMCfg, MistralCfg, VisionCfg, AudioCfg, vidi15_9b
from .engine import Vidi15Engine, make_plan

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200          # vidi/constants.py:10-11
BF16 = torch.bfloat16


@dataclass
class DattnCausalLMOutputWithPast:
    loss: Optional[torch.Tensor] = None
    logits: torch.Tensor = None
    past_key_values: object = None
    past_image_key_values: object = None
    past_audio_key_values: object = None
    hidden_states: object = None
    attentions: object = None


class StreamKVCache:
    """Per-layer K,V of an image/audio stream for the whole batch, exposed like the reference's DynamicCache of 3-D
    ``[B, N, Hkv*dh]`` entries (gemma.py:61-65, 664-670) but backed by the engine's packed per-sample [L, N_b, 2*kv_dim] buffers:
    ``cache[l]`` -> (K [B, Nmax, kv_dim], V [B, Nmax, kv_dim]), samples right-padded with zeros (views when B == 1)."""

    def __init__(self, states: list, which: int, kv_dim: int):
        self.states, self.which, self.kv_dim = states, which, kv_dim

    def _rows(self, b: int, l: int):
        st = self.states[b]
        r0, n = st.seg[self.which][0], st.seg[self.which][1]
        return st.kv[l, r0:r0 + n]

    def __len__(self):
        return self.states[0].kv.shape[0]

    def __getitem__(self, l):
        rows = [self._rows(b, l) for b in range(len(self.states))]
        if len(rows) == 1:
            r = rows[0][None]
        else:
            n = max(x.shape[0] for x in rows)
            r = rows[0].new_zeros(len(rows), n, rows[0].shape[1])
            for b, x in enumerate(rows):
                r[b, :x.shape[0]] = x
        return r[:, :, :self.kv_dim], r[:, :, self.kv_dim:]

    def get_seq_length(self, layer_idx: int = 0):
        return max(st.seg[self.which][1] for st in self.states)


class TextKVCache:
    """The text stream's self-attention cache of every sample (slot ``past_key_values`` of the reference's output, gemma.py:684):
    ``cache[l]`` -> (K [B, Hkv, T, dh], V [B, Hkv, T, dh]) in the HF layout, RoPE already applied to K as HF stores it."""

    def __init__(self, states: list, kv_heads: int, head_dim: int):
        self.states, self.kv_heads, self.head_dim = states, kv_heads, head_dim

    def __len__(self):
        return self.states[0].text_cache["kv"].shape[0]

    def get_seq_length(self, layer_idx: int = 0):
        return max(st.text_cache["len"] for st in self.states)

    def __getitem__(self, l):
        T, kd = self.get_seq_length(), self.kv_heads * self.head_dim
        rows = self.states[0].text_cache["kv"].new_zeros(len(self.states), T, 2 * kd)
        for b, st in enumerate(self.states):
            n = st.text_cache["len"]
            rows[b, :n] = st.text_cache["kv"][l, :n]
        k = rows[:, :, :kd].reshape(len(self.states), T, self.kv_heads, self.head_dim).transpose(1, 2)
        v = rows[:, :, kd:].reshape(len(self.states), T, self.kv_heads, self.head_dim).transpose(1, 2)
        return k, v


class _PrefillState:
    """What `generate` carries between steps (gemma.py:657-687 carries image/audio embeds + 3 caches)."""

    def __init__(self, kv, seg, text_cache):
        self.kv, self.seg, self.text_cache = kv, seg, text_cache


def hf_like_config(cfg) -> SimpleNamespace:
    """``model.config`` as the reference's callers read it (plain attribute bag; mm_splits stays settable, inference.py:86)."""
    c = cfg.llm
    gemma = hasattr(c, "final_softcap")
    # eos: 107 for the Gemma2 build (gemma.py:461-462); Mistral keeps the tokenizer's </s> = 2
    # mm_image_aspect_ratio: read by the reference's process_images(video, image_processor, model.config) (img_utils.py:173-198);
    # anything but "resize" / "pad" / "anyres" / "crop" raises there -- the checkpoints ship "resize" (Dattn*Config defaults)
    return SimpleNamespace(mm_splits=cfg.mm_splits, eos_token_id=107 if gemma else 2, pad_token_id=0,
                           vocab_size=c.vocab, hidden_size=c.hidden, num_hidden_layers=c.layers,
                           mm_image_aspect_ratio=getattr(cfg, "mm_image_aspect_ratio", "resize"),
                           mm_input_type="video", mm_image_pool_size=cfg.mm_image_pool_size,
                           mm_audio_pool_size=cfg.mm_audio_pool_size, mm_time_interval=cfg.mm_time_interval,
                           final_logit_softcapping=getattr(c, "final_softcap", None),
                           model_type="dattn_gemma2" if gemma else "dattn_mistral")


class DattnGemma2ForCausalLM:
    """Inference-only stand-in for the reference class of the same name."""

    def __init__(self, cfg: Vidi15Config, state_dict: dict, device="cuda", tokenizer=None, image_processor=None,
                 audio_processor=None, rank: int = 0, world: int = 1, group=None, pop_state_dict=False):
        self.engine = Vidi15Engine(cfg, state_dict, device=device, rank=rank, world=world, group=group,
                                   pop_state_dict=pop_state_dict)
        self.cfg = cfg
        self.device = self.engine.device
        self.dtype = BF16
        self.config = hf_like_config(cfg)
        self._mm = SimpleNamespace(text_tokenizer=tokenizer, image_processor=image_processor, audio_processor=audio_processor)
        self.training = False

    # --- reference surface -------------------------------------------------------------------------
    def get_model(self):
        return self._mm

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def half(self):
        return self

    def cuda(self):
        return self

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    # --- helpers -----------------------------------------------------------------------------------
    def _dev(self, t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if t is None:
            return None
        return t.to(device=self.device, dtype=BF16, non_blocking=True).contiguous()

    def _any_nonzero(self, shard: torch.Tensor) -> bool:
        """`sum(abs(x)) != 0` over the whole sample (multimodal.py:202,246), evaluated on the uploaded shard and
        OR-ed over ranks.  One tiny device reduction + one host read per modality per forward."""
        flag = shard.any().to(torch.int32) if shard.numel() else torch.zeros((), device=self.device, dtype=torch.int32)
        if self.engine.world > 1:
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MAX, group=self.engine.group)
        return bool(flag.item())

    @staticmethod
    def _strip(ids_row: torch.Tensor, mask_row: Optional[torch.Tensor]) -> torch.Tensor:
        if mask_row is not None:
            ids_row = ids_row[mask_row.bool()]
        n_img = int((ids_row == IMAGE_TOKEN_INDEX).sum())
        assert n_img <= 1, "only support at most one image for now."      # multimodal.py:369
        return ids_row[ids_row != IMAGE_TOKEN_INDEX]

    def _prefill_one(self, ids, images, audios, audio_size, max_len, logits_to_keep, mm_total=None):
        """mm_total=(F_total, C_total): `images` / `audios` already hold only THIS RANK's contiguous shard of frames /
        chunks (engine.make_plan) -- the host-memory-friendly multi-GPU mode; otherwise they are the full tensors and
        the shard is sliced here (the reference replicates inputs inside an SP group, vidi_trainer.py:222-254)."""
        eng = self.engine
        img = aud = None
        F = Cn = 0
        iv = av = True
        if mm_total is not None:
            F, Cn = int(mm_total[0]), int(mm_total[1])
            plan = make_plan(self.cfg, F, Cn, audio_size or 0, eng.rank, eng.world)
            if images is not None:
                assert images.shape[0] == plan.f1 - plan.f0, "images must be this rank's frame shard"
                img = self._dev(images); iv = self._any_nonzero(img)
            if audios is not None:
                assert audios.shape[0] == plan.c1 - plan.c0, "audios must be this rank's chunk shard"
                aud = self._dev(audios); av = self._any_nonzero(aud)
        else:
            if images is not None:
                F = images.shape[0]
                plan = make_plan(self.cfg, F, audios.shape[0] if audios is not None else 0, audio_size or 0, eng.rank, eng.world)
                img = self._dev(images[plan.f0:plan.f1])
                iv = self._any_nonzero(img)                              # multimodal.py:202 (input validity, not compute)
            if audios is not None:
                Cn = audios.shape[0]
                plan = make_plan(self.cfg, F, Cn, audio_size or 0, eng.rank, eng.world)
                aud = self._dev(audios[plan.c0:plan.c1])
                av = self._any_nonzero(aud)
        tc = eng.new_text_cache(max_len)
        ids_dev = ids.to(self.device, dtype=torch.int64).contiguous()
        logits, st = eng.prefill(ids_dev, img, aud, audio_size or 0, n_frames_total=F, n_chunks_total=Cn,
                                 logits_to_keep=logits_to_keep, text_cache=tc, image_valid=iv, audio_valid=av,
                                 return_state=True)
        return logits, _PrefillState(st["kv"], st["seg"], tc)

    @torch.no_grad()
    def encode_media(self, images: Optional[torch.Tensor], audios: Optional[torch.Tensor], audio_size: int = 0):
        """Towers + projectors + the whole image/audio stream pass of ONE video, without any text: returns the state that
        ``generate(..., media=state)`` runs queries against.  Not in the reference (its streams are recomputed for every ask()); the
        decomposed attention makes them query-independent (gemma.py:183-202), which the batched VUE runner exploits."""
        eng = self.engine
        F = images.shape[0] if images is not None else 0
        Cn = audios.shape[0] if audios is not None else 0
        plan = make_plan(self.cfg, F, Cn, audio_size or 0, eng.rank, eng.world)
        img = self._dev(images[plan.f0:plan.f1]) if images is not None else None
        aud = self._dev(audios[plan.c0:plan.c1]) if audios is not None else None
        iv = self._any_nonzero(img) if img is not None else True
        av = self._any_nonzero(aud) if aud is not None else True
        S, seg = eng.encode_streams(img, aud, plan, iv, av)
        return _PrefillState(eng.stream_pass(S), seg, None)

    # --- forward -----------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, input_ids: torch.LongTensor = None, attention_mask: Optional[torch.Tensor] = None,
                position_ids=None, past_key_values=None, past_image_key_values=None, past_audio_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                images: Optional[torch.Tensor] = None, image_sizes=None, image_embeds=None, image_attention_mask=None,
                audios: Optional[torch.Tensor] = None, audio_sizes: Optional[List[int]] = None, audio_embeds=None,
                audio_attention_mask=None, return_dict=None, cache_position=None, logits_to_keep: int = 0, **kw):
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")      # gemma.py:295-296
        if inputs_embeds is not None:
            raise NotImplementedError("`inputs_embeds` is not supported")                        # gemma.py:615-616
        if labels is not None:
            raise NotImplementedError("training loss is out of scope of the inference engine (gemma.py:572-590)")
        if input_ids.dim() == 1:
            input_ids = input_ids[None]
        B = input_ids.shape[0]
        c = self.cfg.llm
        if isinstance(past_key_values, TextKVCache):
            # continuation on the engine's caches (what HF's loop does through prepare_inputs_for_generation, gemma.py:657-672):
            # the new token ids of every sample run the text stream against the three caches of the prefill
            states = past_key_values.states
            assert len(states) == B, "past_key_values holds a different batch size"
            outs = []
            for b in range(B):
                ids = self._strip(input_ids[b], None).to(self.device, dtype=torch.int64).contiguous()
                st = states[b]
                need = st.text_cache["len"] + ids.numel()
                if need > st.text_cache["kv"].shape[1]:                  # grow the text cache (prefill sized it for T + 1)
                    grown = self.engine.new_text_cache(max(need, 2 * st.text_cache["kv"].shape[1]))
                    grown["kv"][:, :st.text_cache["len"]] = st.text_cache["kv"][:, :st.text_cache["len"]]
                    grown["len"] = st.text_cache["len"]
                    st.text_cache = grown
                outs.append(self.engine.text_pass(ids, st.kv, st.seg, text_cache=st.text_cache, logits_to_keep=logits_to_keep))
        else:
            outs, states = [], []
            for b in range(B):
                ids = self._strip(input_ids[b], attention_mask[b] if attention_mask is not None else None)   # None mask == ones (Q16)
                img = images[b] if images is not None else None
                aud = audios[b] if audios is not None else None
                asz = int(audio_sizes[b]) if audio_sizes is not None else (aud.shape[0] * self.cfg.aud.nb_max_frames if aud is not None else 0)
                lg, st = self._prefill_one(ids, img, aud, asz, ids.numel() + 1, logits_to_keep, mm_total=kw.get("mm_total"))
                outs.append(lg); states.append(st)
        T = max(o.shape[0] for o in outs)
        logits = torch.zeros(B, T, outs[0].shape[1], device=self.device, dtype=torch.float32)
        for b, o in enumerate(outs):
            logits[b, :o.shape[0]] = o                                    # right padding (gemma.py:459 padding_side)
        # the three cache slots hold EVERY sample of the batch (gemma.py:664-670, 684-685)
        kinds = [k for k, present in (("img", images is not None or isinstance(past_image_key_values, StreamKVCache)),
                                      ("aud", audios is not None or isinstance(past_audio_key_values, StreamKVCache))) if present]
        if isinstance(past_key_values, TextKVCache):
            img_c, aud_c = past_image_key_values, past_audio_key_values
        else:
            img_c = StreamKVCache(states, kinds.index("img"), c.kv_dim) if "img" in kinds else None
            aud_c = StreamKVCache(states, kinds.index("aud"), c.kv_dim) if "aud" in kinds else None
        return DattnCausalLMOutputWithPast(logits=logits, past_key_values=TextKVCache(states, c.kv_heads, c.head_dim),
                                           past_image_key_values=img_c, past_audio_key_values=aud_c)

    # --- generate ----------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, inputs: Optional[torch.Tensor] = None, images: Optional[torch.Tensor] = None, image_sizes=None,
                 audios: Optional[torch.Tensor] = None, audio_sizes: Optional[List[int]] = None, media=None, **kwargs) -> torch.LongTensor:
        """Greedy decoding (the only mode the reference's callers use: do_sample=False, inference.py:40-50).
        Returns only the new token ids, as HF does when generation starts from embeddings (gemma.py:646-655)."""
        if "inputs_embeds" in kwargs:
            raise NotImplementedError("`inputs_embeds` is not supported")                        # gemma.py:615-616
        if kwargs.get("do_sample", False):
            raise NotImplementedError("only greedy decoding (do_sample=False) is implemented")
        max_new = int(kwargs.get("max_new_tokens", 1024))
        eos = kwargs.get("eos_token_id", self.config.eos_token_id)
        eos = set(eos) if isinstance(eos, (list, tuple)) else {eos}
        pad = kwargs.get("pad_token_id", None)
        pad = 0 if pad is None else pad
        attention_mask = kwargs.get("attention_mask", None)
        if inputs.dim() == 1:
            inputs = inputs[None]
        B = inputs.shape[0]
        seqs = []
        for b in range(B):
            ids = self._strip(inputs[b], attention_mask[b] if attention_mask is not None else None)
            img = images[b] if images is not None else None
            aud = audios[b] if audios is not None else None
            asz = int(audio_sizes[b]) if audio_sizes is not None else (aud.shape[0] * self.cfg.aud.nb_max_frames if aud is not None else 0)
            if media is not None:            # streams of this video already encoded (encode_media): only the text pass runs
                assert B == 1 and images is None and audios is None, "media= carries one video's streams; pass no images / audios"
                tc = self.engine.new_text_cache(ids.numel() + max_new)
                logits = self.engine.text_pass(ids.to(self.device, dtype=torch.int64).contiguous(), media.kv, media.seg, text_cache=tc,
                                               logits_to_keep=1)
                st = _PrefillState(media.kv, media.seg, tc)
            else:
                logits, st = self._prefill_one(ids, img, aud, asz, ids.numel() + max_new, logits_to_keep=1)
            new = []
            for _ in range(max_new):
                nxt = int(torch.argmax(logits[-1]))
                new.append(nxt)
                if nxt in eos or len(new) == max_new:
                    break
                step_ids = torch.tensor([nxt], device=self.device, dtype=torch.int64)
                logits = self.engine.text_pass(step_ids, st.kv, st.seg, text_cache=st.text_cache, logits_to_keep=1)
            seqs.append(new)
        n = max(len(s) for s in seqs)
        out = torch.full((B, n), pad, dtype=torch.long)
        for b, s in enumerate(seqs):
            out[b, :len(s)] = torch.tensor(s, dtype=torch.long)
        return out.to(self.device)


class DattnMistralForCausalLM(DattnGemma2ForCausalLM):
    """Vidi-7B twin (Vidi_7B/model/lmm/dattn/mistral.py:496-713): same surface, Mistral-family engine paths
    (selected by the config type: SwiGLU, no post-norms / soft-caps / normaliser, learned-conv pooling, fp32 logits)."""


# -------------------------------------------------------------------------------------------------
# loader (builder.py:24-64)
# -------------------------------------------------------------------------------------------------
def config_from_hf_json(cfg_json: dict):
    """Build the engine config from a checkpoint's config.json: keys of DattnGemma2Config (gemma.py:427-448) or, for
    ``model_type == "dattn_mistral"`` / a Mistral architecture, DattnMistralConfig (Vidi_7B/model/lmm/dattn/mistral.py:456-477),
    the dispatch ``get_dattn_cls`` makes on the checkpoint (Vidi_7B/model/builder.py:49)."""
    g = cfg_json.get
    # tower dims: the public SigLIP-so400m/14@384 and Whisper-large-v3 configs unless the checkpoint carries overrides
    vis = VisionCfg(**g("vision_config")) if isinstance(g("vision_config"), dict) else VisionCfg()
    aud = AudioCfg(**g("audio_config")) if isinstance(g("audio_config"), dict) else AudioCfg()
    archs = " ".join(g("architectures") or [])
    mm = dict(mm_audio_pool_size=g("mm_audio_pool_size") or 5, mm_time_interval=g("mm_time_interval") or 10000,
              mm_std=g("mm_std") or 0.028976401314139366, mm_splits=g("mm_splits") or 1)
    aspect = g("mm_image_aspect_ratio") or "resize"
    if g("model_type") == "dattn_mistral" or "Mistral" in archs:       # mm_image_aspect_ratio is carried into model.config (img_utils.py:173-198)
        hidden, heads = g("hidden_size", 4096), g("num_attention_heads", 32)
        llm = MistralCfg(hidden=hidden, heads=heads, kv_heads=g("num_key_value_heads", 8), head_dim=g("head_dim") or hidden // heads,
                         inter=g("intermediate_size", 14336), layers=g("num_hidden_layers", 32), vocab=g("vocab_size", 32000),
                         rms_eps=g("rms_norm_eps", 1e-5), rope_theta=g("rope_theta", 10000.0),
                         tie_word_embeddings=g("tie_word_embeddings", False), sliding_window=g("sliding_window") or 0)
        return Vidi7BConfig(llm=llm, vis=vis, aud=aud, mm_image_pool_size=g("mm_image_pool_size") or 16, mm_image_aspect_ratio=aspect, **mm)
    else:
        llm = LLMCfg(hidden=g("hidden_size", 3584), heads=g("num_attention_heads", 16), kv_heads=g("num_key_value_heads", 8),
                     head_dim=g("head_dim", 256), inter=g("intermediate_size", 14336), layers=g("num_hidden_layers", 42),
                     vocab=g("vocab_size", 256000), rms_eps=g("rms_norm_eps", 1e-6), rope_theta=g("rope_theta", 10000.0),
                     query_pre_attn_scalar=g("query_pre_attn_scalar", 256), attn_softcap=g("attn_logit_softcapping", 50.0),
                     final_softcap=g("final_logit_softcapping", 30.0), sliding_window=g("sliding_window", 4096),
                     tie_word_embeddings=g("tie_word_embeddings", True))
    return Vidi15Config(llm=llm, vis=vis, aud=aud, mm_image_pool_size=g("mm_image_pool_size") or 2, mm_image_aspect_ratio=aspect, **mm)


def _load_safetensors_dir(path: str) -> dict:
    from safetensors.torch import load_file
    sd = {}
    files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    for f in files:
        sd.update(load_file(os.path.join(path, f)))
    return sd


def load_pretrained_model(model_name_or_path, load_8bit=False, load_4bit=False, device_map="auto", device="cuda",
                          use_flash_attn=True, **kwargs):
    """-> (model, tokenizer, image_processor, audio_processor), as builder.py:24-64.

    ``model_name_or_path`` is a directory with ``config.json`` + HF-layout ``*.safetensors`` shards (the key layout of
    SURVEY.md 8b).  8/4-bit loading is not part of the prefill path and raises.  ``use_flash_attn`` / ``device_map`` are
    accepted for signature compatibility; attention always runs on the engine's own kernels on one device per rank."""
    if load_8bit or load_4bit:
        raise NotImplementedError("8-bit / 4-bit loading (builder.py:30-40) is out of scope of the B200 prefill engine")
    with open(os.path.join(model_name_or_path, "config.json")) as f:
        cfg = config_from_hf_json(json.load(f))
    sd = _load_safetensors_dir(model_name_or_path)
    tokenizer = image_processor = audio_processor = None
    try:                                                                     # present only when the ckpt dir ships them
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(model_name_or_path, model_max_length=4096, padding_side="right")
    except Exception:                                                        # noqa: BLE001
        pass
    from .preprocess import SiglipImageProcessorLite, WhisperFeatureExtractorLite
    image_processor = SiglipImageProcessorLite(cfg.vis.image)
    audio_processor = WhisperFeatureExtractorLite(cfg.aud.mels)
    cls = DattnMistralForCausalLM if isinstance(cfg, Vidi7BConfig) else DattnGemma2ForCausalLM      # get_dattn_cls (builder.py:49)
    if isinstance(cfg, Vidi7BConfig) and tokenizer is not None and getattr(tokenizer, "pad_token", None) is None:
        tokenizer.pad_token = tokenizer.unk_token                       # DattnMistralMMModel.build_text_tokenizer (mistral.py:487-493)
    model = cls(cfg, sd, device=device, tokenizer=tokenizer, image_processor=image_processor, audio_processor=audio_processor,
                pop_state_dict=True)
    return model, tokenizer, image_processor, audio_processor

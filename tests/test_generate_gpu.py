"""Decoded output parity (north_star: "decoded time-range indices bit-exact"): generate() / ask() on the engine vs the fp32 oracle on a
checkpoint whose greedy decode spells a real time-range answer with decisive margins (tests/chain_fixture.py), plus teacher-forced
decode steps through forward(past_key_values=...) against the oracle's re-prefill logits, and per-sample caches for B > 1."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import chain_fixture as CF  # noqa: E402

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _fixture(question="a dog running."):
    from vidi_b200 import pipeline as P
    from vidi_b200.config import vidi15_mini
    from vidi_b200.model import DattnGemma2ForCausalLM
    cfg = vidi15_mini()
    tok = CF.CharTokenizer(cfg.llm.vocab)
    ids = P.build_input_ids(question, tok, "vidi15")
    cfg, sd = CF.make_chain_checkpoint(cfg, int(ids[0, -1]), gain=2.0)
    sd = {k: (v if "mm_rand_pos" in k else v.to(BF).float()) for k, v in sd.items()}
    model = DattnGemma2ForCausalLM(cfg, {k: v.clone() for k, v in sd.items()}, device="cuda", tokenizer=tok)
    return cfg, sd, tok, ids, model


def test_generate_ids_text_and_time_ranges_equal_oracle():
    """Full equality of the generated ids (no soft matching), of the decoded text and of the formatted time ranges; the oracle's top-2
    margin at every step must exceed 1.0 logit (the engine's logit error is ~0.05), so equality is a fair demand."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.postprocess import format_time_ranges
    cfg, sd, tok, ids, model = _fixture()
    _, images, mels, asz = synth.make_inputs(cfg, 4, 1, n_text=5)
    images, mels = images.to(BF).float(), mels.to(BF).float()
    ref_ids, margins = R.greedy_generate(sd, cfg, ids[0], images, mels, asz, max_new_tokens=40, normalizer_dtype=BF, return_margins=True)
    assert min(margins) > 1.0 and ref_ids[-1] == tok.eos_token_id and ref_ids[:-1] == CF.answer_ids()
    gen = model.generate(ids, images=images[None], audios=mels[None], audio_sizes=[asz], do_sample=False, max_new_tokens=40,
                         use_cache=True, disable_compile=True, pad_token_id=tok.pad_token_id)
    assert gen[0].tolist() == ref_ids
    text = tok.batch_decode(gen, skip_special_tokens=True)[0].strip()
    assert text == CF.ANSWER == tok.batch_decode([ref_ids])[0].strip()
    assert format_time_ranges(text, 4000.0) == "00:06:40-00:16:40, 00:33:20-00:50:00"


def test_ask_end_to_end_device_preprocessing_equals_oracle_pipeline():
    """pipeline.ask (inference.py:18-66) on the REAL engine from decoded media that already sits on the GPU: uint8 frames -> resize kernels,
    16 kHz samples -> log-mel kernels + GEMMs, prompt + sentinel, generate, regex, timestamps.  The same call with host media (torch
    restatement of Pillow / the HF extractor) and the oracle as the model must give the same string."""
    from oracle import vidi15_ref as R
    from vidi_b200 import pipeline as P
    from vidi_b200.preprocess import SiglipImageProcessorLite, WhisperFeatureExtractorLite
    cfg, sd, tok, ids, model = _fixture("a dog running.")
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (5, 180, 320, 3), generator=g, dtype=torch.uint8)
    t = torch.arange(16000 * 7) / 16000.0
    audio = 0.2 * torch.sin(2 * torch.pi * 330 * t) + 0.05 * torch.randn(t.numel(), generator=g)
    ip, ap = SiglipImageProcessorLite(cfg.vis.image), WhisperFeatureExtractorLite(cfg.aud.mels)
    length = 4000.0
    out = P.ask("a dog running.", frames.cuda(), audio.cuda(), length, model, tok, ip, ap, max_new_tokens=40)

    class OracleModel:            # the oracle behind generate(): host fp32 media, bf16-rounded like the engine's inputs
        def generate(self, input_ids, images=None, audios=None, audio_sizes=None, max_new_tokens=40, **kw):
            new = R.greedy_generate(sd, cfg, input_ids[0], images[0].to(BF).float(), audios[0].to(BF).float(), audio_sizes[0],
                                    max_new_tokens=max_new_tokens, normalizer_dtype=BF)
            return torch.tensor([new])
    ref = P.ask("a dog running.", frames, audio, length, OracleModel(), tok, ip, ap, max_new_tokens=40)
    assert out == ref == "00:06:40-00:16:40, 00:33:20-00:50:00"
    # the device pre-processing itself: bit-equal frames, log-mel within 1e-3 (+ bf16 rounding)
    assert torch.equal(ip.preprocess(frames.cuda()).cpu(), ip.preprocess(frames).to(BF))
    mel_d, n_d = ap(audio.cuda()); mel_h, n_h = ap(audio)
    assert n_d == n_h and float((mel_d.float().cpu() - mel_h).abs().max()) <= 1e-3 + 2 ** -8 * float(mel_h.abs().max())


def test_forward_continuation_teacher_forced_matches_oracle_prefill():
    """Decode through the reference-shaped surface: forward(prompt, images=, audios=) then forward(new ids, past_key_values=...,
    past_image_key_values=..., past_audio_key_values=...) with ARBITRARY forced tokens (not the argmax, so a wrong cache position or
    window cannot hide behind a repeated token).  Each step's logits vs the oracle's full re-prefill of the extended text."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.config import vidi15_mini
    from vidi_b200.model import DattnGemma2ForCausalLM, TextKVCache
    cfg = vidi15_mini()
    sd = {k: (v if "mm_rand_pos" in k else v.to(BF).float()) for k, v in synth.make_state_dict(cfg, seed=21).items()}
    model = DattnGemma2ForCausalLM(cfg, {k: v.clone() for k, v in sd.items()}, device="cuda")
    ids, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=12, seed=8, audio_size=900)
    images, mels = images.to(BF).float(), mels.to(BF).float()
    out = model(ids[None], images=images[None], audios=mels[None], audio_sizes=[asz])
    assert isinstance(out.past_key_values, TextKVCache) and out.past_key_values.get_seq_length() == 12
    cur = ids.clone()
    for step, forced in enumerate([[17], [900, 3], [511]]):
        new = torch.tensor(forced)
        out = model(new[None], past_key_values=out.past_key_values, past_image_key_values=out.past_image_key_values,
                    past_audio_key_values=out.past_audio_key_values)
        cur = torch.cat([cur, new])
        ref = R.prefill(sd, cfg, cur, images, mels, asz, normalizer_dtype=BF)[-len(forced):]
        assert out.logits.shape == (1, len(forced), cfg.llm.vocab)
        assert rel(out.logits[0], ref) < 3e-2, (step, rel(out.logits[0], ref))
        assert float((out.logits[0].cpu() - ref).abs().max()) < 0.25
    k, v = out.past_key_values[0]
    assert k.shape == (1, cfg.llm.kv_heads, 12 + 4, cfg.llm.head_dim)


def test_forward_batch_returns_caches_of_every_sample():
    """B = 2: the three cache slots hold both samples (gemma.py:664-670, 684-685), right-padded like the logits."""
    from oracle import synth, vidi15_ref as R
    from vidi_b200.config import vidi15_mini
    from vidi_b200.model import DattnGemma2ForCausalLM
    cfg = vidi15_mini()
    sd = {k: (v if "mm_rand_pos" in k else v.to(BF).float()) for k, v in synth.make_state_dict(cfg, seed=11).items()}
    model = DattnGemma2ForCausalLM(cfg, {k: v.clone() for k, v in sd.items()}, device="cuda")
    ids0, img0, mel0, a0 = synth.make_inputs(cfg, 2, 1, n_text=9, seed=1, audio_size=500)
    ids1, img1, mel1, a1 = synth.make_inputs(cfg, 2, 1, n_text=5, seed=2, audio_size=900)
    img0, img1, mel0, mel1 = [t.to(BF).float() for t in (img0, img1, mel0, mel1)]
    L = max(ids0.numel(), ids1.numel())
    ids = torch.zeros(2, L, dtype=torch.long); mask = torch.zeros(2, L, dtype=torch.long)
    ids[0, :ids0.numel()] = ids0; mask[0, :ids0.numel()] = 1
    ids[1, :ids1.numel()] = ids1; mask[1, :ids1.numel()] = 1
    out = model(ids, attention_mask=mask, images=torch.stack([img0, img1]), audios=torch.stack([mel0, mel1]), audio_sizes=[a0, a1])
    na = [cfg.audio_tokens(a0), cfg.audio_tokens(a1)]
    k_a, v_a = out.past_audio_key_values[1]
    assert k_a.shape == (2, max(na), cfg.llm.kv_dim) and float(k_a[0, na[0]:].abs().max()) == 0.0
    k_i, _ = out.past_image_key_values[0]
    assert k_i.shape == (2, cfg.image_tokens(2), cfg.llm.kv_dim)
    for b, (i_, im, me, a_) in enumerate([(ids0, img0, mel0, a0), (ids1, img1, mel1, a1)]):
        _, inter = R.prefill(sd, cfg, i_, im, me, a_, normalizer_dtype=BF, return_intermediates=True)
        assert rel(k_i[b], inter["kv"][0][0][0]) < 2e-2                      # layer-0 image K of sample b
        assert rel(k_a[b, :na[b]], inter["kv"][1][1][0]) < 3e-2              # layer-1 audio K of sample b
    kt, vt = out.past_key_values[0]
    assert kt.shape == (2, cfg.llm.kv_heads, 9, cfg.llm.head_dim) and float(kt[1, :, 5:].abs().max()) == 0.0


def test_vidi7b_loader_generate_and_ask(tmp_path):
    """Vidi-7B boundary (Vidi_7B/model/builder.py:25-65, Vidi_7B/inference.py:19-65): a checkpoint directory whose config.json says
    model_type dattn_mistral loads into DattnMistralForCausalLM (learned-conv pool weights included); its logits equal the directly
    constructed model; greedy ids, decoded text and the formatted time range equal the Mistral-family oracle on the chain fixture."""
    import dataclasses
    import json
    from safetensors.torch import save_file
    from oracle import synth, vidi7b_ref as R7
    from vidi_b200 import pipeline as P
    from vidi_b200.config import Vidi7BConfig, vidi7b_mini
    from vidi_b200.model import DattnMistralForCausalLM, load_pretrained_model
    from vidi_b200.preprocess import SiglipImageProcessorLite, WhisperFeatureExtractorLite
    cfg = vidi7b_mini()
    tok = CF.CharTokenizer(cfg.llm.vocab, "mistral")
    length = 4000.0
    ids = P.build_input_ids("a dog running.", tok, "vidi7b", length_s=length)
    cfg, sd = CF.make_chain_checkpoint(cfg, int(ids[0, -1]), gain=1.0, eos_id=tok.eos_token_id, answer="0.10-0.25", embed_scale=20.0)
    sd_bf = {k: (v if "mm_rand_pos" in k else v.to(BF)) for k, v in sd.items()}
    sd = {k: v.float() for k, v in sd_bf.items()}
    keys = sorted(sd_bf)
    save_file({k: sd_bf[k].contiguous() for k in keys[::2]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: sd_bf[k].contiguous() for k in keys[1::2]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    c = cfg.llm
    hf = dict(model_type="dattn_mistral", architectures=["DattnMistralForCausalLM"], hidden_size=c.hidden, num_attention_heads=c.heads,
              num_key_value_heads=c.kv_heads, head_dim=c.head_dim, intermediate_size=c.inter, num_hidden_layers=c.layers, vocab_size=c.vocab,
              rms_norm_eps=c.rms_eps, rope_theta=c.rope_theta, tie_word_embeddings=False, sliding_window=None,
              mm_image_pool_size=cfg.mm_image_pool_size, mm_audio_pool_size=5, mm_time_interval=10000, mm_std=cfg.mm_std,
              mm_image_aspect_ratio="resize", mm_input_type="video", vision_config=dataclasses.asdict(cfg.vis),
              audio_config=dataclasses.asdict(cfg.aud))
    (tmp_path / "config.json").write_text(json.dumps(hf))
    model, _, img_proc, aud_proc = load_pretrained_model(str(tmp_path))
    assert isinstance(model, DattnMistralForCausalLM) and isinstance(model.cfg, Vidi7BConfig) and not model.engine.gemma
    assert model.cfg.llm.head_dim == c.head_dim and model.config.mm_image_aspect_ratio == "resize" and model.config.model_type == "dattn_mistral"
    direct = DattnMistralForCausalLM(cfg, {k: v.clone() for k, v in sd.items()}, device="cuda")
    _, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=5, audio_size=1500)
    images, mels = images.to(BF).float(), mels.to(BF).float()
    a = model(ids, images=images[None], audios=mels[None], audio_sizes=[asz]).logits
    b = direct(ids, images=images[None], audios=mels[None], audio_sizes=[asz]).logits
    assert torch.equal(a, b) and a.dtype == torch.float32                        # logits.float() (mistral.py:615-616)
    ref = R7.prefill(sd, cfg, ids[0], images, mels, asz)
    assert rel(a[0], ref) < 3e-2
    ref_ids, margins = R7.greedy_generate(sd, cfg, ids[0], images, mels, asz, max_new_tokens=16, eos_id=tok.eos_token_id, return_margins=True)
    assert min(margins) > 1.0 and ref_ids[-1] == tok.eos_token_id
    gen = model.generate(ids, images=images[None], audios=mels[None], audio_sizes=[asz], do_sample=False, max_new_tokens=16,
                         use_cache=True, pad_token_id=tok.pad_token_id)
    assert gen[0].tolist() == ref_ids
    assert tok.batch_decode(gen)[0] == "0.10-0.25"
    # ask(), Vidi-7B prompt (video length inside the question), device-side pre-processing
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (3, 120, 200, 3), generator=g, dtype=torch.uint8)
    audio = 0.1 * torch.randn(16000 * 5, generator=g)
    ip, ap = SiglipImageProcessorLite(cfg.vis.image), WhisperFeatureExtractorLite(cfg.aud.mels)
    out = P.ask("a dog running.", frames.cuda(), audio.cuda(), length, model, tok, ip, ap, family="vidi7b", max_new_tokens=16)
    assert out == "00:06:40-00:16:40"


def test_vue_runner_on_engine_encodes_each_video_once():
    """vue_runner.run_queries on the real engine: two queries on one video and one on another; every record carries the chain fixture's
    answer scaled by its duration; the towers + stream pass run once per video (launch counter), every query only the text stream."""
    from vidi_b200 import ops, vue_runner as V
    from vidi_b200.preprocess import SiglipImageProcessorLite, WhisperFeatureExtractorLite
    cfg, sd, tok, ids, model = _fixture()
    g = torch.Generator().manual_seed(9)
    vids = {v: (torch.randint(0, 256, (3 + i, 90, 160, 3), generator=g, dtype=torch.uint8).cuda(), (0.1 * torch.randn(16000 * 4, generator=g)).cuda())
            for i, v in enumerate(("vA", "vB"))}
    qs = [dict(query_id=1, video_id="vA", duration=4000.0, query="a dog running."),
          dict(query_id=2, video_id="vB", duration=200.0, query="somebody opens a door"),
          dict(query_id=5, video_id="vA", duration=4000.0, query="a red car")]
    ip, ap = SiglipImageProcessorLite(cfg.vis.image), WhisperFeatureExtractorLite(cfg.aud.mels)
    encodes = []
    orig = model.encode_media
    model.encode_media = lambda *a, **k: (encodes.append(1), orig(*a, **k))[1]
    recs = V.run_queries(qs, lambda v: vids[v], model, tok, ip, ap, max_new_tokens=40)
    assert len(encodes) == 2
    assert [r["query_id"] for r in recs] == [1, 2, 5]
    assert recs[0]["answer"] == recs[2]["answer"] == [[0.10 * 4000.0, 0.25 * 4000.0], [0.50 * 4000.0, 0.75 * 4000.0]]
    assert recs[1]["answer"] == [[0.10 * 200.0, 0.25 * 200.0], [0.50 * 200.0, 0.75 * 200.0]]
    # same answer as the one-shot path (generate with images / audios)
    video = ip.preprocess(vids["vB"][0]); feats, asz = ap(vids["vB"][1])
    from vidi_b200 import pipeline as P
    one = model.generate(P.build_input_ids(qs[1]["query"], tok, "vidi15"), images=video[None], audios=feats[None], audio_sizes=[asz],
                         do_sample=False, max_new_tokens=40)
    assert tok.batch_decode(one)[0] == CF.ANSWER

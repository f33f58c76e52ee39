"""Pins the oracle (oracle/vidi15_ref.py) against fixtures produced by running the REFERENCE'S OWN modules and the HF
blocks it subclasses (tests/golden/make_golden.py, generated in the build container where /root/reference exists)."""
import os

import pytest
import torch

from oracle import vidi15_ref as R
from vidi_b200 import synth
from vidi_b200.config import AudioCfg, LLMCfg, Vidi15Config, VisionCfg

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vidi15_reference_golden.pt"), weights_only=False)


def close(a, b, tol=2e-5):
    a, b = a.float(), b.float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = float((a - b).abs().max())
    assert err <= tol * max(1.0, float(b.abs().max())), err


def tiny_cfg():
    c = G["cfg"]
    return Vidi15Config(llm=LLMCfg(**c["llm"]), vis=VisionCfg(**c["vis"]), aud=AudioCfg(**c["aud"]), name="golden-tiny")


def test_conv2dpool_and_space_to_depth_match_reference():
    for key, ref in G["pool"]["out"].items():
        h, w = map(int, key.split("x"))
        close(R.conv2d_pool(G["pool"]["x"], (h, w), 2), ref, 1e-6)
    close(R.space_to_depth(G["s2d"]["x"], 2), G["s2d"]["out"], 0)


def test_resize_by_tokens_matches_reference():
    cfg = Vidi15Config()
    for B, hw in G["resize_by_tokens"].items():
        side = cfg.vis.side + 1
        if B * side * side > cfg.max_image_tokens * 4:
            assert cfg.image_hw(B) == tuple(hw), (B, cfg.image_hw(B), hw)
    assert cfg.image_hw(306) == (28, 28) and cfg.image_hw(3600) == (10, 10) and cfg.image_hw(600) == (20, 20)


def test_pos_embed_norm_mlp_match_reference():
    p = G["pos"]
    sd = {f"h.{k}": v for k, v in p["h_sd"].items()} | {f"t.{k}": v for k, v in p["t_sd"].items()}
    close(R.pos_embed(sd, "h", 14, 2, 32), p["h_out"])
    close(R.pos_embed(sd, "t", 37, 10000, 32), p["t_out"], 2e-4)     # sin/cos of arguments up to 1e4 rad
    n = G["norm"]
    close(R.mm_norm(n["x"], n["w"], 1e-5), n["out"], 1e-6)
    close(R.xhat(n["x"], 1e-5), n["out_plain"], 1e-6)
    m = G["mlp"]
    close(R.projector({f"p.{k}": v for k, v in m["sd"].items()}, "p", m["x"]), m["out"], 1e-6)


def test_towers_match_installed_hf_blocks():
    cfg = tiny_cfg()
    sd = synth.make_state_dict(cfg, seed=G["seed"])
    ids, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=6, seed=99, audio_size=1234)
    close(R.siglip_tower(sd, cfg, images), G["siglip"]["hidden_m2"], 2e-5)
    close(R.whisper_encoder(sd, cfg, mels), G["whisper"]["out"], 2e-5)


def test_encode_video_matches_reference_mixin():
    cfg = tiny_cfg()
    sd = synth.make_state_dict(cfg, seed=G["seed"])
    ids, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=6, seed=99, audio_size=1234)
    e = G["encode"]
    X, mX = R.encode_video_images(sd, cfg, images)
    A, mA = R.encode_video_audios(sd, cfg, mels, asz)
    close(X, e["image_embeds"], 5e-5); close(A, e["audio_embeds"], 5e-5)
    assert torch.equal(mX, e["image_mask"].bool()) and torch.equal(mA, e["audio_mask"].bool())
    assert X.shape[0] == cfg.image_tokens(3) and A.shape[0] == cfg.audio_tokens(asz)


def test_decoder_layers_match_reference_layer_forward():
    cfg = tiny_cfg()
    sd = synth.make_state_dict(cfg, seed=G["seed"])
    d = G["decoder"]
    H, S_img, S_aud = d["H0"], d["img0"], d["aud0"]
    T = H.shape[0]
    cos, sin = R.rope_cos_sin(T, cfg.llm.head_dim, cfg.llm.rope_theta)
    ones_i, ones_a = torch.ones(S_img.shape[0], dtype=torch.bool), torch.ones(S_aud.shape[0], dtype=torch.bool)
    for l, ref in enumerate(d["layers"]):
        p = f"model.layers.{l}"
        S_img2, Ki, Vi = R.stream_layer(S_img, sd, p, cfg)
        S_aud2, Ka, Va = R.stream_layer(S_aud, sd, p, cfg)
        H = R.text_layer(H, sd, p, cfg, l, cos, sin, [(Ki, Vi, ones_i), (Ka, Va, ones_a)])
        S_img, S_aud = S_img2, S_aud2
        close(S_img, ref["image"], 5e-5); close(S_aud, ref["audio"], 5e-5); close(H, ref["text"], 5e-5)


def test_prefill_consistency_with_layer_fixture():
    """The oracle's own prefill() must reproduce the layer-by-layer path that was pinned above."""
    cfg = tiny_cfg()
    sd = synth.make_state_dict(cfg, seed=G["seed"])
    ids, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=6, seed=99, audio_size=1234)
    logits, inter = R.prefill(sd, cfg, ids, images, mels, asz, return_intermediates=True)
    close(inter["text_hidden"][-1], G["decoder"]["layers"][-1]["text"], 5e-5)
    assert logits.shape == (6, cfg.llm.vocab) and float(logits.abs().max()) <= 30.0


def test_prefill_logits_match_reference_model_forward():
    """End to end against the reference's own DattnGemma2ForCausalLM.forward -> prepare_inputs_labels_for_multimodal ->
    DattnGemma2Model.forward (gemma.py:267-424,484-601; multimodal.py:339-451) run unmodified on the same seeded tiny model:
    sentinel stripping, embedding, normaliser, the 3-stream layer loop, final norm, lm_head and the 30*tanh(x/30) soft-cap."""
    cfg = tiny_cfg()
    sd = synth.make_state_dict(cfg, seed=G["seed"])
    ids, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=6, seed=99, audio_size=1234)
    logits, inter = R.prefill(sd, cfg, ids, images, mels, asz, return_intermediates=True)
    close(logits, G["model"]["logits"], 5e-5)
    assert torch.equal(logits.argmax(-1), G["model"]["logits"].argmax(-1))
    if "final_hidden" in inter:
        close(inter["final_hidden"], G["model"]["last_hidden_state"], 5e-5)


def test_vidi7b_conv_pool_matches_reference():
    """Vidi_7B/model/mm_vision/pool.py Conv2DPool (learned conv + align_corners bilinear), run from the reference file."""
    from oracle import vidi7b_ref as R7
    g = G["pool7b"]
    for s_out, case in g["cases"].items():
        close(R7.conv2d_pool_7b(g["x"], case["w"], s_out), case["out"], 1e-6)


def test_vidi7b_oracle_runs_and_shapes():
    from oracle import vidi7b_ref as R7
    from vidi_b200.config import vidi7b_mini
    cfg = vidi7b_mini()
    sd = synth.make_state_dict(cfg, seed=5)
    ids, images, mels, asz = synth.make_inputs(cfg, 2, 1, n_text=7, audio_size=900)
    logits = R7.prefill(sd, cfg, ids, images, mels, asz)
    assert logits.shape == (7, cfg.llm.vocab) and torch.isfinite(logits).all()


def test_vidi7b_decoder_layers_match_reference_layer_forward():
    """oracle/vidi7b_ref.py stream_layer + text_layer vs the reference's own DattnMistralDecoderLayer.forward
    (tests/golden/make_golden_7b.py: two stacked layers, all three streams, varlen cross-attention path)."""
    from oracle import vidi7b_ref as R7
    from vidi_b200.config import MistralCfg, Vidi7BConfig
    G7 = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vidi7b_reference_golden.pt"), weights_only=False)
    c = G7["cfg"]
    cfg = Vidi7BConfig(llm=MistralCfg(**c["llm"]), vis=VisionCfg(**c["vis"]), aud=AudioCfg(**c["aud"]), mm_image_pool_size=c["pool"])
    sd = synth.make_state_dict(cfg, seed=G7["seed"])
    H, S_img, S_aud = G7["H0"], G7["img0"], G7["aud0"]
    cos, sin = R.rope_cos_sin(H.shape[0], cfg.llm.head_dim, cfg.llm.rope_theta)
    ones_i, ones_a = torch.ones(S_img.shape[0], dtype=torch.bool), torch.ones(S_aud.shape[0], dtype=torch.bool)
    for l, ref in enumerate(G7["layers"]):
        p = f"model.layers.{l}"
        S_img2, Ki, Vi = R7.stream_layer(S_img, sd, p, cfg)
        S_aud2, Ka, Va = R7.stream_layer(S_aud, sd, p, cfg)
        H = R7.text_layer(H, sd, p, cfg, cos, sin, [(Ki, Vi, ones_i), (Ka, Va, ones_a)])
        S_img, S_aud = S_img2, S_aud2
        close(S_img, ref["image"], 5e-5); close(S_aud, ref["audio"], 5e-5); close(H, ref["text"], 5e-5)


def test_vidi7b_prefill_matches_reference_model_forward():
    """Vidi-7B end to end against the reference's own DattnMistralForCausalLM.forward -> prepare_inputs_labels_for_multimodal /
    encode_video_images / encode_video_audios (Vidi_7B/model/lmm/dattn/multimodal.py:154-227) -> DattnMistralModel.forward
    (mistral.py:296-453,512-616), run unmodified (tests/golden/make_golden_7b.py): encoder composition, learned-conv pooling,
    input prep, layer loop, final norm, untied lm_head with fp32 logits."""
    from oracle import vidi7b_ref as R7
    from vidi_b200.config import MistralCfg, Vidi7BConfig
    G7 = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vidi7b_reference_golden.pt"), weights_only=False)
    c, g = G7["cfg"], G7["model"]
    cfg = Vidi7BConfig(llm=MistralCfg(**c["llm"]), vis=VisionCfg(**c["vis"]), aud=AudioCfg(**c["aud"]), mm_image_pool_size=c["pool"])
    sd = synth.make_state_dict(cfg, seed=G7["seed"])
    ids, images, mels, asz = synth.make_inputs(cfg, 3, 1, n_text=6, seed=98, audio_size=g["audio_size"])
    logits, inter = R7.prefill(sd, cfg, ids, images, mels, asz, return_intermediates=True)
    close(inter["image_embeds"], g["image_embeds"], 5e-5)
    close(inter["audio_embeds"], g["audio_embeds"], 5e-5)
    close(logits, g["logits"], 5e-5)
    assert torch.equal(logits.argmax(-1), g["logits"].argmax(-1))

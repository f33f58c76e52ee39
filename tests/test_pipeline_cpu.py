"""Caller-side helpers: prompt / sentinel construction against the reference's own txt_utils functions
(tests/golden/make_golden_text.py runs them unmodified on the deterministic FakeTokenizer), and the ask() plumbing end to end with a
stub model."""
import json
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from fake_tokenizer import FakeTokenizer  # noqa: E402

from vidi_b200 import pipeline as P  # noqa: E402
from vidi_b200.preprocess import SiglipImageProcessorLite, WhisperFeatureExtractorLite  # noqa: E402

G = json.load(open(os.path.join(HERE, "golden", "text_utils_golden.json")))


@pytest.mark.parametrize("family,tok_family", [("vidi15", "gemma2"), ("vidi7b", "mistral")])
@pytest.mark.parametrize("add_bos", [True, False])
def test_tokenizer_image_token_and_chat_prompt_match_reference(family, tok_family, add_bos):
    tok = FakeTokenizer(tok_family, add_bos)
    g = G[family][f"bos{int(add_bos)}"]
    for prompt, ref in zip(G["prompts"], g["ids"]):
        assert P.tokenizer_image_token(prompt, tok) == ref
        assert P.tokenizer_image_token(prompt, tok, return_tensors="pt").tolist() == ref
    for q, ref in zip(G["questions"], g["chat"]):
        assert P.chat_prompt([{"from": "human", "value": "<image>\n" + q}], tok, family) == ref
    with pytest.raises(ValueError):
        P.tokenizer_image_token("x", tok, return_tensors="np")


def test_build_input_ids_has_one_sentinel_and_strips_period():
    tok = FakeTokenizer("gemma2")
    ids = P.build_input_ids("a man opening a door.", tok, "vidi15")
    assert ids.shape[0] == 1 and ids.dtype == torch.long and int((ids == P.IMAGE_TOKEN_INDEX).sum()) == 1 and int(ids[0, 0]) == tok.bos_token_id
    assert torch.equal(ids, P.build_input_ids("a man opening a door", tok, "vidi15"))
    ids7 = P.build_input_ids("two cats.", FakeTokenizer("mistral"), "vidi7b", length_s=93.5)
    assert int((ids7 == P.IMAGE_TOKEN_INDEX).sum()) == 1


class _StubModel:
    """records what ask() hands to generate() and answers with a fixed id sequence"""
    def __init__(self, answer_ids):
        self.answer, self.calls = answer_ids, []

    def generate(self, input_ids, **kw):
        self.calls.append((input_ids, kw))
        return torch.tensor([self.answer])


def test_ask_plumbing_with_stub_model():
    tok = FakeTokenizer("gemma2")
    answer = "0.10-0.25, 0.50-0.75"
    answer_ids = tok(answer).input_ids[1:]
    model = _StubModel(answer_ids)
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (4, 72, 128, 3), generator=g, dtype=torch.uint8)
    audio = 0.1 * torch.randn(16000 * 4, generator=g)
    ip, ap = SiglipImageProcessorLite(64), WhisperFeatureExtractorLite(128)
    out = P.ask("a dog running.", frames, audio, 4000.0, model, tok, ip, ap)
    (ids, kw), = model.calls
    assert ids.shape[0] == 1 and int((ids == -200).sum()) == 1
    assert kw["images"].shape == (1, 4, 3, 64, 64) and kw["audios"].shape == (1, 1, 128, 3000) and kw["audio_sizes"] == [400]
    assert kw["do_sample"] is False and kw["max_new_tokens"] == 1024 and kw["use_cache"] is True and kw["disable_compile"] is True
    # the fake decode joins tokens with spaces ("0 . 10 - 0 . 25 ..."): compare against the regex-based formatter on the plain text
    from vidi_b200.postprocess import format_time_ranges
    assert format_time_ranges(answer, 4000.0) == "00:06:40-00:16:40, 00:33:20-00:50:00"
    assert isinstance(out, str)


def test_vue_runner_schema_grouping_and_scorer_join(tmp_path):
    """Batched VUE-TR-V2 runner (vue_runner.py): result records in the README.md:78-95 schema, one media encode per VIDEO (not per query),
    and the records survive the join qa_eval.py::load_result performs (query_id join, floor / ceil of the predicted ranges)."""
    from vidi_b200 import vue_runner as V
    tok = FakeTokenizer("gemma2")

    class Stub:
        def __init__(self):
            self.encodes, self.generates = 0, 0

        def encode_media(self, video, feats, audio_size):
            self.encodes += 1
            return ("media", video.shape[0], audio_size)

        def generate(self, ids, media=None, **kw):
            self.generates += 1
            assert media[0] == "media" and kw["do_sample"] is False and int((ids == -200).sum()) == 1
            return torch.tensor([[1, 2, 3]])

    tok.batch_decode = lambda ids, skip_special_tokens=True: ["0.10-0.25, 0.5-0.75"]
    model = Stub()
    gts = [dict(query_id=7, video_id="vA", duration=4000.0, query="a dog.", gt=[[405, 990]], task="temporal_retrieval"),
           dict(query_id=3, video_id="vB", duration=100.0, query="a cat", gt=[[0, 10]], task="temporal_retrieval"),
           dict(query_id=9, video_id="vA", duration=4000.0, query="two dogs", gt=[[2000, 3000]], task="temporal_retrieval")]
    g = torch.Generator().manual_seed(0)

    def media(vid):
        return torch.randint(0, 256, (3, 36, 64, 3), generator=g, dtype=torch.uint8), 0.1 * torch.randn(16000 * 3, generator=g)
    recs = V.run_queries(gts, media, model, tok, SiglipImageProcessorLite(32), WhisperFeatureExtractorLite(128))
    assert model.encodes == 2 and model.generates == 3                      # vA encoded once for its two queries
    assert [r["query_id"] for r in recs] == [7, 3, 9]
    assert set(recs[0]) == {"query_id", "video_id", "duration", "query", "answer", "task"}
    assert recs[0]["answer"] == [[400.0, 1000.0], [2000.0, 3000.0]] and recs[1]["answer"] == [[10.0, 25.0], [50.0, 75.0]]
    path = tmp_path / "results_vidi_b200.json"
    V.write_results(str(path), recs)
    rows = V.merge_with_ground_truth(gts, json.load(open(path)))
    assert rows[0]["gt"] == [[405, 990]] and rows[0]["answer"][0] == [400, 1000]
    assert abs(V.temporal_iou(rows[0]["answer"][:1], rows[0]["gt"]) - 585 / 600) < 1e-9
    assert V.temporal_iou(rows[2]["answer"][1:], rows[2]["gt"]) == 1.0


def test_media_decoding_follows_load_video_and_load_audio(tmp_path):
    """media.load_video / load_audio / get_length (vid_utils.py:9-49, inference.py:68-75): a clip written here with OpenCV (10 fps, 37
    frames, a distinct grey level per frame) is sampled at every round(avg_fps / fps)-th frame from 0; time_range uses linspace over
    the frame-index range; a 16 kHz 16-bit .wav decodes to float32 / 32768 mono."""
    import wave
    import numpy as np
    cv2 = pytest.importorskip("cv2")
    from vidi_b200 import media as M
    path = str(tmp_path / "clip.mp4")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (64, 48))
    if not wr.isOpened():
        pytest.skip("OpenCV build cannot write mp4v")
    for i in range(37):
        wr.write(np.full((48, 64, 3), 5 + 6 * i, np.uint8))
    wr.release()
    frames = M.load_video(path, fps=1.0)
    assert frames.dtype == torch.uint8 and frames.shape == (4, 48, 64, 3)                 # frames 0, 10, 20, 30
    levels = frames.float().mean((1, 2, 3))
    assert torch.allclose(levels, torch.tensor([5.0, 65.0, 125.0, 185.0]), atol=4.0)     # lossy codec: right frames, approximate values
    sub = M.load_video(path, fps=2.0, time_range=(1.0, 3.0))                              # 4 frames over indices 10..30
    assert sub.shape[0] == 4 and abs(float(sub[0].float().mean()) - 65.0) < 4 and abs(float(sub[-1].float().mean()) - 185.0) < 4
    assert abs(M.get_length(path) - 3.7) < 0.11
    wav = str(tmp_path / "a.wav")
    pcm = (np.sin(np.arange(16000 * 2) * 0.05) * 12000).astype(np.int16)
    with wave.open(wav, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
    audio = M.load_audio(wav, 16000)
    assert audio.dtype == torch.float32 and audio.shape == (32000,) and torch.allclose(audio, torch.from_numpy(pcm.astype(np.float32) / 32768.0))
    assert WhisperFeatureExtractorLite(128).audio_size(audio.numel()) == 200
    # ask_path end to end with a stub model (host pre-processing)
    tok = FakeTokenizer("gemma2")
    tok.batch_decode = lambda ids, skip_special_tokens=True: ["0.10-0.50"]
    out = M.ask_path("a grey ramp.", path, _StubModel([1, 2, 3]), tok, SiglipImageProcessorLite(32), WhisperFeatureExtractorLite(128),
                     audio_path=wav, device=None)
    assert out == "00:00:00-00:00:01"

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

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

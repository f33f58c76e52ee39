"""A synthetic checkpoint whose greedy decode SPELLS a known time-range answer, plus the character tokenizer that reads it.

No real weights or tokenizer are reachable offline, and a random-init model with tied embeddings just repeats one token, which makes
"decoded time ranges bit-exact" (north_star) untestable.  Construction (legit under DattnGemma2Config: ``tie_word_embeddings=False``):
the residual stream at a position is dominated by that position's token embedding, so an output head whose row ``b`` equals the
embedding row of token ``a`` makes ``b`` the greedy successor of ``a``.  Chaining   last prompt token -> a_0 -> a_1 -> ... -> eos
with one distinct token id per answer position makes generate() emit the answer's characters in order; every other head row stays
random.  The oracle (fp32) and the engine (bf16) must then produce the SAME ids, the same decoded text and the same formatted
"HH:MM:SS-HH:MM:SS" string; the oracle also reports the top-2 margin of every step so the tests can assert the fixture is decisive.
Test infrastructure only."""
import dataclasses
import re
from types import SimpleNamespace

import torch

ANSWER = "0.10-0.25, 0.50-0.75"
CHARS = "0123456789.-, "            # char of token id i (i >= BASE) is CHARS[(i - BASE) % len(CHARS)]
BASE = 300


class CharTokenizer:
    """Word-level ids for the prompt (hashed into [108, BASE), clear of bos / pad / eos), character tokens for ids >= BASE; the
    Gemma-2 chat template with eos 107 (gemma.py:461-462), or the Mistral-instruct one with bos 1 / eos 2."""

    def __init__(self, vocab: int, family: str = "gemma2"):
        self.vocab, self.family = vocab, family
        if family == "gemma2":
            self.bos_token, self.bos_token_id, self.pad_token_id, self.eos_token_id = "<bos>", 2, 0, 107
        else:
            self.bos_token, self.bos_token_id, self.pad_token_id, self.eos_token_id = "<s>", 1, 0, 2

    def __call__(self, text):
        import zlib
        words = re.findall(r"<[a-z_/]+>|\[/?INST\]|\w+|[^\w\s]|\n", text)
        return SimpleNamespace(input_ids=[self.bos_token_id] + [108 + zlib.crc32(w.encode()) % (BASE - 108) for w in words])

    def apply_chat_template(self, messages, tokenize=False):
        out = self.bos_token
        if self.family != "gemma2":
            for m in messages:
                out += ("[INST] " + m["content"] + " [/INST]") if m["role"] == "user" else (m["content"] + "</s>")
            return out
        for m in messages:
            out += "<start_of_turn>" + ("model" if m["role"] == "assistant" else m["role"]) + "\n" + m["content"].strip() + "<end_of_turn>\n"
        return out

    def batch_decode(self, ids, skip_special_tokens=True):
        rows = []
        for row in ids:
            s = ""
            for i in row:
                i = int(i)
                if i >= BASE:
                    s += CHARS[(i - BASE) % len(CHARS)]
                elif not skip_special_tokens:
                    s += f"<{i}>"
            rows.append(s)
        return rows


def answer_ids(answer: str = ANSWER):
    """one distinct id per answer position, spelling the answer under CharTokenizer"""
    n = len(CHARS)
    return [BASE + n * k + CHARS.index(ch) for k, ch in enumerate(answer)]


def make_chain_checkpoint(cfg, last_prompt_id: int, seed: int = 1234, answer: str = ANSWER, gain: float = 1.0, eos_id: int = 107,
                          embed_scale: float = 1.0):
    """-> (cfg with untied head, state_dict fp32).  The caller rounds to bf16 as usual.  embed_scale: the Mistral family has no
    sqrt(D) embedding normaliser (gemma.py:353), so its fixture scales the embedding table instead to make the token identity
    dominate the residual stream."""
    from vidi_b200 import synth
    cfg = dataclasses.replace(cfg, llm=dataclasses.replace(cfg.llm, tie_word_embeddings=False))
    sd = synth.make_state_dict(cfg, seed=seed)
    sd["model.embed_tokens.weight"] = sd["model.embed_tokens.weight"] * embed_scale
    E, H = sd["model.embed_tokens.weight"], sd["lm_head.weight"]
    ids = answer_ids(answer)
    assert max(ids) < cfg.llm.vocab, "vocab too small for the answer"
    prev = last_prompt_id
    for t in ids + [eos_id]:
        H[t] = gain * E[prev] / embed_scale
        prev = t
    return cfg, sd

"""A deterministic stand-in for the HF tokenizer (none is available offline) used by the text-utility fixtures and tests: word-level
ids from a hash, BOS prepended to every call like the Gemma / Mistral tokenizers do, and the two chat templates of the model cards
(Gemma-2: <bos><start_of_turn>role\\ncontent<end_of_turn>\\n ; Mistral-instruct: <s>[INST] content [/INST])."""
import re
import zlib
from types import SimpleNamespace


class FakeTokenizer:
    def __init__(self, family="gemma2", add_bos=True):
        self.family, self.add_bos = family, add_bos
        self.bos_token = "<bos>" if family == "gemma2" else "<s>"
        self.bos_token_id, self.pad_token_id = 2, 0
        self._vocab = {}

    def _id(self, w):
        return 10 + zlib.crc32(w.encode()) % 50000

    def __call__(self, text):
        words = re.findall(r"<[a-z_/]+>|\[/?INST\]|\w+|[^\w\s]|\n", text)
        ids = [self._id(w) for w in words]
        for w, i in zip(words, ids):
            self._vocab[i] = w
        return SimpleNamespace(input_ids=([self.bos_token_id] if self.add_bos else []) + ids)

    def apply_chat_template(self, messages, tokenize=False):
        assert not tokenize
        if self.family == "gemma2":
            out = self.bos_token
            for m in messages:
                role = "model" if m["role"] == "assistant" else m["role"]
                out += "<start_of_turn>" + role + "\n" + m["content"].strip() + "<end_of_turn>\n"
            return out
        out = self.bos_token
        for m in messages:
            out += ("[INST] " + m["content"] + " [/INST]") if m["role"] == "user" else (m["content"] + "</s>")
        return out

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(self._vocab.get(int(i), "?") for i in row) for row in ids]

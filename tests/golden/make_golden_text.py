"""tests/golden/text_utils_golden.json: outputs of the REFERENCE's own text utilities — tokenizer_image_token / preprocess_chat of
Vidi1.5_9B/vidi/dataset/txt_utils.py and Vidi_7B/model/txt_utils.py, imported unmodified through ref_shim — on the deterministic
FakeTokenizer (no HF tokenizer is available offline).  Run in the build container: python tests/golden/make_golden_text.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

ref_shim.install()
from fake_tokenizer import FakeTokenizer  # noqa: E402
from vidi.dataset import txt_utils as T15  # noqa: E402  (reference code)

PROMPTS = ["<image>\nDuring which time segments in the video can we see a dog?", "no image here", "<image>", "a<image>b<image>c",
           "<image>\nwhat, exactly; happens?\n<image> and then"]
QUESTIONS = ["a man opening a door.", "two cats", "slicing onions in a kitchen."]
out = dict(prompts=PROMPTS, questions=QUESTIONS, vidi15={}, vidi7b={})
for add_bos in (True, False):
    tok = FakeTokenizer("gemma2", add_bos)
    key = f"bos{int(add_bos)}"
    out["vidi15"][key] = dict(
        ids=[T15.tokenizer_image_token(p, tok) for p in PROMPTS],
        chat=[T15.preprocess_chat([{"from": "human", "value": "<image>\n" + q}], tok) for q in QUESTIONS])

# the 7B tree is a top-level package called `model`
import types  # noqa: E402

_pkg = types.ModuleType("model")                      # skip Vidi_7B/model/__init__.py (it imports the flash-attn-2 Mistral classes)
_pkg.__path__ = ["/root/reference/Vidi_7B/model"]
sys.modules["model"] = _pkg
import model.txt_utils as T7  # noqa: E402  (reference code, loaded as a submodule of the stub package)

for add_bos in (True, False):
    tok = FakeTokenizer("mistral", add_bos)
    key = f"bos{int(add_bos)}"
    out["vidi7b"][key] = dict(
        ids=[T7.tokenizer_image_token(p, tok) for p in PROMPTS],
        chat=[T7.preprocess_chat([{"from": "human", "value": "<image>\n" + q}], tok) for q in QUESTIONS])
with open(os.path.join(HERE, "text_utils_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote text_utils_golden.json")

"""The caller side of the prefill path, mirroring ``ask()`` (Vidi1.5_9B/vidi/eval/inference.py:18-66, Vidi_7B/inference.py:19-65)
from DECODED inputs: prompt construction, the ``<image>`` sentinel, pre-processing, ``model.generate`` and the timestamp
post-processing.  Video / audio decoding (decord, ffmpeg) stays outside — pass uint8 RGB frames sampled at 1 fps and 16 kHz mono
float32 samples.  SURVEY.md 8a rows a1, a5."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .postprocess import format_time_ranges

IMAGE_TOKEN_INDEX = -200                 # vidi/constants.py
DEFAULT_IMAGE_TOKEN = "<image>"
PROMPT_VIDI15 = "During which time segments in the video can we see {}?"                       # inference.py:34
PROMPT_VIDI7B = ("Given the frames from a video, answer the time range in percentage that corresponds to query text split by "
                 "comma. Video length is: {:.2f} and text query is: {}.")                       # Vidi_7B/inference.py:34


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX, return_tensors: Optional[str] = None):
    """txt_utils.py:14-33: tokenize the text around every ``<image>`` separately and join the pieces with the sentinel id; a BOS
    that the tokenizer prepends to each piece is kept once, at the front."""
    pieces = [tokenizer(chunk).input_ids for chunk in prompt.split(DEFAULT_IMAGE_TOKEN)]
    has_bos = bool(pieces) and len(pieces[0]) > 0 and pieces[0][0] == tokenizer.bos_token_id
    ids: List[int] = [pieces[0][0]] if has_bos else []
    skip = 1 if has_bos else 0
    for i, piece in enumerate(pieces):
        if i > 0:
            ids.append(image_token_index)
        ids.extend(piece[skip:])
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError(f"Unsupported tensor type: {return_tensors}")


def chat_prompt(source: Sequence[Dict[str, str]], tokenizer, family: str = "vidi15") -> str:
    """``preprocess_chat`` (txt_utils.py:148-155 / Vidi_7B/model/txt_utils.py:135-140): the tokenizer's own chat template over
    alternating human/gpt turns with the BOS string removed; the Gemma2 family appends the open model turn."""
    roles_chat = ("user", "model") if family == "vidi15" else ("user", "assistant")
    messages = []
    for i, turn in enumerate(source):
        assert turn["from"] == ("human", "gpt")[i % 2]
        messages.append({"role": roles_chat[i % 2], "content": turn["value"]})
    conversation = tokenizer.apply_chat_template(messages, tokenize=False)
    if tokenizer.bos_token:
        conversation = conversation.replace(tokenizer.bos_token, "")
    if family == "vidi15":
        conversation += "<start_of_turn>model\n"
    return conversation


def build_input_ids(question: str, tokenizer, family: str = "vidi15", length_s: float = 0.0) -> torch.Tensor:
    """inference.py:33-37: strip one trailing period, wrap in the family's fixed prompt behind ``<image>\\n``, chat-template it and
    tokenize around the sentinel -> [1, T+1] int64."""
    q = question[:-1] if question.endswith(".") else question
    text = PROMPT_VIDI15.format(q) if family == "vidi15" else PROMPT_VIDI7B.format(length_s, q)
    prompt = chat_prompt([{"from": "human", "value": DEFAULT_IMAGE_TOKEN + "\n" + text}], tokenizer, family)
    return tokenizer_image_token(prompt, tokenizer, IMAGE_TOKEN_INDEX, return_tensors="pt").unsqueeze(0)


def ask(question: str, frames_uint8: torch.Tensor, audio: torch.Tensor, length_s: float, model, tokenizer, image_processor,
        audio_processor, family: str = "vidi15", max_new_tokens: int = 1024) -> str:
    """``ask()`` from decoded media: frames [F,H,W,3] uint8 (1 fps), audio [n] float32 mono 16 kHz, length_s = media length in
    seconds -> "HH:MM:SS-HH:MM:SS, ..." (inference.py:18-66)."""
    video = image_processor.preprocess(frames_uint8)                       # process_images, 'resize' branch
    audio_feats, audio_size = audio_processor(audio)                       # process_audio
    input_ids = build_input_ids(question, tokenizer, family, length_s)
    kw = dict(images=video.unsqueeze(0), audios=audio_feats.unsqueeze(0), audio_sizes=[audio_size], do_sample=False,
              max_new_tokens=max_new_tokens, use_cache=True, pad_token_id=tokenizer.pad_token_id)
    if family == "vidi15":
        kw["disable_compile"] = True
    with torch.inference_mode():
        output_ids = model.generate(input_ids, **kw)
    text = tokenizer.batch_decode(output_ids, skip_special_tokens=True)[0].strip()
    return format_time_ranges(text, length_s)

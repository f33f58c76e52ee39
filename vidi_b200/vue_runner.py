"""Batched temporal-retrieval runner in the VUE-TR-V2 result schema.

The reference evaluates by calling ``ask()`` once per (video, query) pair and writing a list of
``{"query_id", "video_id", "duration", "query", "answer": [[start_s, end_s], ...], "task"}`` records that
``VUE_TR_V2/qa_eval.py::load_result`` (lines 303-337) joins with the ground truth by ``query_id`` (README.md:78-95).  Here the
queries are grouped by video: the image / audio streams never read the text, so the towers and the whole
stream pass run ONCE per video (``model.encode_media``) and every query of that video only pays the text pass and the decode.
Decoding of the media itself stays outside (pass decoded frames / samples through ``media_provider``)."""
from __future__ import annotations

import json
from collections import OrderedDict
from typing import Callable, Dict, Iterable, List, Tuple

import torch

from .pipeline import build_input_ids
from .postprocess import parse_ranges


def answers_in_seconds(text: str, duration: float) -> List[List[float]]:
    """generated text -> [[start_s, end_s], ...]: the normalised ``a-b`` fractions of inference.py:55-58 scaled by the duration"""
    return [[a * duration, b * duration] for a, b in parse_ranges(text)]


def run_queries(queries: Iterable[dict], media_provider: Callable[[str], Tuple[torch.Tensor, torch.Tensor]], model, tokenizer,
                image_processor, audio_processor, family: str = "vidi15", max_new_tokens: int = 1024) -> List[dict]:
    """queries: dicts with query_id, video_id, duration, query (and optionally task) -- the ground-truth file's records work as they
    are.  media_provider(video_id) -> (frames uint8 [F,H,W,3] sampled at 1 fps, audio float32 [n] mono 16 kHz), on the CPU or the GPU.
    Returns the records of the result file, in the order of ``queries``."""
    by_video: "OrderedDict[str, List[dict]]" = OrderedDict()
    for q in queries:
        by_video.setdefault(q["video_id"], []).append(q)
    out: Dict[object, dict] = {}
    for vid, qs in by_video.items():
        frames, audio = media_provider(vid)
        video = image_processor.preprocess(frames)
        feats, audio_size = audio_processor(audio)
        media = model.encode_media(video, feats, audio_size)              # towers + stream pass, once per video
        for q in qs:
            ids = build_input_ids(q["query"], tokenizer, family, float(q["duration"]))
            gen = model.generate(ids, media=media, do_sample=False, max_new_tokens=max_new_tokens, pad_token_id=tokenizer.pad_token_id)
            text = tokenizer.batch_decode(gen, skip_special_tokens=True)[0].strip()
            out[q["query_id"]] = dict(query_id=q["query_id"], video_id=vid, duration=q["duration"], query=q["query"],
                                      answer=answers_in_seconds(text, float(q["duration"])), task=q.get("task", "temporal_retrieval"))
    return [out[q["query_id"]] for q in queries]


def write_results(path: str, records: List[dict]) -> None:
    with open(path, "w") as f:
        json.dump(records, f, indent=1)


def merge_with_ground_truth(gts: List[dict], preds: List[dict]) -> List[dict]:
    """What qa_eval.py::load_result does before scoring: join by query_id, floor the starts and ceil the ends of the predictions."""
    import math
    gt_by_id = {g["query_id"]: g for g in gts}
    rows = []
    for p in preds:
        qid = p["query_id"] if "query_id" in p else p["id"]
        ans = [[math.floor(a), math.ceil(b)] for a, b in p["answer"] if True] if p["answer"] and p["answer"] != [[]] else []
        rows.append({**p, **gt_by_id[qid], "answer": ans})
    return rows


def temporal_iou(pred: List[List[float]], gt: List[List[float]]) -> float:
    """union-of-intervals IoU on the 1-second grid the scorer uses"""
    def cover(rs):
        s = set()
        for a, b in rs:
            s.update(range(int(a), int(b)))
        return s
    p, g = cover(pred), cover(gt)
    return len(p & g) / len(p | g) if (p or g) else 1.0

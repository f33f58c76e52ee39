"""VUE-TR-V2 metrics: the numbers `VUE_TR_V2/qa_eval.py` prints for a result file -- IoU success-rate AUC
(`success_overlap`, qa_eval.py:140-153 on `overlap_ratio` :105-137), and precision / recall threshold-curve AUCs
(`compute_precision_recall`, :262-300) -- after the join `load_result` performs (:303-340: floor the predicted starts, ceil the ends).
Plots, radar charts and the pandas table are not reproduced; the attribute breakdown is `score(records, attribute=(key, value))`.
Pinned against the reference's own functions on seeded synthetic result sets (tests/golden/make_golden_vue.py -> vue_scores_golden.json)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_THRES = np.linspace(0, 1, 101)
_trapz = getattr(np, "trapezoid", None) or np.trapz            # numpy >= 2 renamed trapz


def join_with_ground_truth(gts: Sequence[dict], preds: Sequence[dict]) -> List[dict]:
    """load_result: join by query_id (or id); answers become [[floor(start), ceil(end)], ...]; [] and [[]] mean "no prediction" """
    by_id = {g["query_id"]: g for g in gts}
    rows = []
    for p in preds:
        qid = p["query_id"] if "query_id" in p else p["id"]
        a = p["answer"]
        ans = [] if (len(a) == 0 or (len(a) == 1 and len(a[0]) == 0)) else [[math.floor(x[0]), math.ceil(x[1])] for x in a]
        row = {**p, **by_id[qid]}
        row["answer"] = ans
        rows.append(row)
    return rows


def _merge(intervals: np.ndarray) -> np.ndarray:
    """merge_time_spans: sort by start, merge overlapping or touching spans"""
    iv = intervals[np.argsort(intervals[:, 0], kind="stable")].astype(float)
    out = [iv[0].copy()]
    for s, e in iv[1:]:
        if s <= out[-1][1]:
            out[-1][1] = max(out[-1][1], e)
        else:
            out.append(np.array([s, e]))
    return np.array(out)


def iou(pred: Sequence[Sequence[float]], gt: Sequence[Sequence[float]]) -> float:
    """overlap_ratio: predictions are merged first, inverted spans dropped; intersection summed over all (pred, gt) pairs;
    union = len(pred) + len(gt) - intersection; clipped to [0, 1]; both empty -> 1, one empty -> 0."""
    g = np.array(gt, dtype=float).reshape(-1, 2) if len(gt) else np.zeros((0, 2))
    p = np.array(pred, dtype=float).reshape(-1, 2) if len(pred) else np.zeros((0, 2))
    if g.shape[0] == 0:
        return 1.0 if p.shape[0] == 0 else 0.0
    if p.shape[0] == 0:
        return 0.0
    p = _merge(p)
    len_gt = float(np.sum(g[:, 1] - g[:, 0]))
    p = p[p[:, 0] <= p[:, 1]]
    inter = 0.0
    for ps, pe in p:
        for gs, ge in g:
            inter += max(0.0, min(pe, ge) - max(ps, gs))
    union = float(np.sum(p[:, 1] - p[:, 0])) + len_gt - inter
    return float(min(1.0, max(0.0, inter / (union + 1e-16))))


def iou_auc(rows: Sequence[dict]) -> Tuple[np.ndarray, float]:
    """success_overlap: share of queries with IoU > t for t in linspace(0, 1, 101), and its trapezoid AUC"""
    v = np.array([iou(r["answer"], r["gt"]) for r in rows])
    success = np.array([np.sum(v > t) / float(len(rows) + 1e-16) for t in _THRES])
    return success, float(_trapz(success, _THRES))


def _intersection(a: List[List[float]], b: List[List[float]]) -> List[Tuple[float, float]]:
    i = j = 0
    out = []
    while i < len(a) and j < len(b):
        (as_, ae), (bs, be) = a[i], b[j]
        if as_ <= be and bs <= ae:
            out.append((max(as_, bs), min(ae, be)))
        if ae < be:
            i += 1
        else:
            j += 1
    return out


def precision_recall(rows: Sequence[dict], avg: bool = True):
    """compute_precision_recall: per query, spans are normalised to [min, max] but NOT merged or sorted (the reference's two-pointer
    intersection runs on them as given); recall is defined where the ground truth is non-empty, precision where the prediction is
    non-empty (or both are empty: 1); avg=True returns the AUCs of the share-of-queries-above-threshold curves."""
    gt_len, pr_len, in_len = [], [], []
    for r in rows:
        gt = [[min(x), max(x)] for x in r["gt"] if len(x) == 2]
        pr = [[min(x), max(x)] for x in r["answer"] if len(x) == 2]
        inter = _intersection([list(x) for x in gt], [list(x) for x in pr])
        gt_len.append(sum(e - s for s, e in gt)); pr_len.append(sum(e - s for s, e in pr)); in_len.append(sum(e - s for s, e in inter))
    recall = np.array([i / g for i, g in zip(in_len, gt_len) if g != 0])
    precision = np.array([1.0 if (g == 0 and p == 0) else i / p for i, g, p in zip(in_len, gt_len, pr_len) if (g == 0 and p == 0) or p != 0])
    if not avg:
        return precision, recall
    pt = np.array([np.mean(precision >= t) for t in _THRES])
    rt = np.array([np.mean(recall >= t) for t in _THRES])
    return float(_trapz(pt, _THRES)), float(_trapz(rt, _THRES))


def score(rows: Sequence[dict], attribute: Optional[Tuple[str, str]] = None) -> Dict[str, float]:
    """the three numbers of `print_result` (fractions, not percent) for all rows or for one attribute value, e.g.
    ("duration_category", "ultra-long"), ("query_format", "phrase"), ("query_modality", "audio")"""
    if attribute is not None:
        rows = [r for r in rows if r.get(attribute[0]) == attribute[1]]
    p, r = precision_recall(rows)
    return dict(n=len(rows), precision=p, recall=r, iou=iou_auc(rows)[1])

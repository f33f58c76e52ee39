"""vidi_b200.vue_score against the reference's own scorer (VUE_TR_V2/qa_eval.py: load_result, overlap_ratio, success_overlap,
compute_precision_recall) on seeded synthetic result sets: tests/golden/make_golden_vue.py ran the reference functions unmodified and
committed their outputs; the sets are re-created here from the same seeds."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_vue import synthetic_sets  # noqa: E402

from vidi_b200 import vue_score as S  # noqa: E402

G = json.load(open(os.path.join(HERE, "golden", "vue_scores_golden.json")))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_scores_match_reference_scorer(seed):
    gts, preds = synthetic_sets(seed)
    rows = S.join_with_ground_truth(gts, preds)
    g = G[str(seed)]
    assert len(rows) == g["n"]
    for r, want in zip(rows[:40], g["first_ious"]):
        assert abs(S.iou(r["answer"], r["gt"]) - want) < 1e-12
    got = S.score(rows)
    for k in ("precision", "recall", "iou"):
        assert abs(got[k] - g[k]) < 1e-12, (k, got[k], g[k])
    for name, want in g["by_attribute"].items():
        key, val = name.split("=")
        sub = S.score(rows, attribute=(key, val))
        assert sub["n"] == want["n"]
        for k in ("precision", "recall", "iou"):
            assert abs(sub[k] - want[k]) < 1e-12, (name, k)


def test_edge_cases_of_the_join_and_iou():
    rows = S.join_with_ground_truth([dict(query_id=0, gt=[[10, 20]]), dict(query_id=1, gt=[[5, 6]])],
                                    [dict(id=1, answer=[[]]), dict(query_id=0, answer=[[10.9, 19.1], [19.0, 25.2]])])
    assert rows[0]["answer"] == [] and rows[1]["answer"] == [[10, 20], [19, 26]]
    assert S.iou([], []) == 1.0 and S.iou([[1, 2]], []) == 0.0 and S.iou([], [[1, 2]]) == 0.0
    assert abs(S.iou(rows[1]["answer"], rows[1]["gt"]) - 10 / 16) < 1e-12          # merged prediction [10, 26] vs [10, 20]
    assert S.iou([[30, 20]], [[10, 40]]) == 0.0                                     # inverted span is dropped

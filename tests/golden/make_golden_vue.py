"""Generates tests/golden/vue_scores_golden.json: the REFERENCE's own scorer functions (VUE_TR_V2/qa_eval.py, imported unmodified from
/root/reference with matplotlib stubbed out -- it is only used for plots) run on seeded synthetic ground-truth / prediction sets.
The sets themselves are re-created from the seed by tests/test_vue_score_cpu.py (synthetic_sets below), so no reference data is copied.
    python tests/golden/make_golden_vue.py"""
import importlib.util
import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def synthetic_sets(seed: int, n: int = 240):
    """ground truth records + prediction records in the result-file schema: jittered, missing, empty ([] and [[]]), inverted,
    overlapping, touching and far-off predictions; 1-3 ground-truth spans per query; the three attribute columns of the GT file."""
    rnd = random.Random(seed)
    gts, preds = [], []
    for q in range(n):
        dur = rnd.uniform(20, 4000)
        spans, t = [], rnd.uniform(0, dur * 0.3)
        for _ in range(rnd.choice([1, 1, 2, 3])):
            a = t + rnd.uniform(0, dur * 0.15)
            b = a + rnd.uniform(1, dur * 0.12)
            spans.append([int(a), int(b) + 1])
            t = b + 1
        gts.append(dict(query_id=q, video_id=f"v{q % 37}", duration=dur, query=f"q{q}", gt=spans, task="temporal_retrieval",
                        duration_category=rnd.choice(["ultra-short", "short", "medium", "long", "ultra-long"]),
                        query_format=rnd.choice(["keyword", "phrase", "sentence"]), query_modality=rnd.choice(["audio", "vision", "vision+audio"])))
        kind = rnd.random()
        if kind < 0.08:
            ans = []
        elif kind < 0.12:
            ans = [[]]
        elif kind < 0.22:
            a = rnd.uniform(0, dur); ans = [[a, a + rnd.uniform(0.5, 30)]]                  # somewhere else
        else:
            ans = []
            for s, e in spans:
                if rnd.random() < 0.85:
                    ans.append([s + rnd.uniform(-8, 8), e + rnd.uniform(-8, 8)])
            if rnd.random() < 0.2 and ans:
                ans.append([ans[-1][1] - 2.5, ans[-1][1] + rnd.uniform(1, 20)])            # overlaps the previous span
            if rnd.random() < 0.1 and ans:
                ans.append([ans[0][1], ans[0][1] + 5.0])                                     # touches
            if rnd.random() < 0.06 and ans:
                ans[0] = [ans[0][1], ans[0][0]]                                              # inverted
            rnd.shuffle(ans)
        preds.append(dict(query_id=q, video_id=f"v{q % 37}", duration=dur, query=f"q{q}", answer=ans, task="temporal_retrieval"))
    rnd.shuffle(preds)
    return gts, preds


def main():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("qa_eval", "/root/reference/VUE_TR_V2/qa_eval.py")
    qa = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(qa)
    out = {}
    tmp = os.path.join(HERE, "_tmp_vue")
    os.makedirs(tmp, exist_ok=True)
    for seed in (1, 2, 3):
        gts, preds = synthetic_sets(seed)
        gp, pp = os.path.join(tmp, "gt.json"), os.path.join(tmp, "pred.json")
        json.dump(gts, open(gp, "w")); json.dump(preds, open(pp, "w"))
        rows = qa.load_result(gp, pp)
        _, iou = qa.success_overlap(rows)
        pre, rec = qa.compute_precision_recall(rows)
        rec_ = dict(n=len(rows), precision=float(pre), recall=float(rec), iou=float(iou), by_attribute={})
        for key, vals in (("duration_category", ["ultra-short", "long"]), ("query_format", ["phrase"]), ("query_modality", ["audio", "vision+audio"])):
            for v in vals:
                sub = [r for r in rows if r[key] == v]
                p2, r2 = qa.compute_precision_recall(sub)
                rec_["by_attribute"][f"{key}={v}"] = dict(n=len(sub), precision=float(p2), recall=float(r2), iou=float(qa.success_overlap(sub)[1]))
        rec_["first_ious"] = [float(qa.overlap_ratio(__import__("numpy").array(r["answer"]), r["gt"])) for r in rows[:40]]
        out[str(seed)] = rec_
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)
    json.dump(out, open(os.path.join(HERE, "vue_scores_golden.json"), "w"), indent=1)
    print({k: {m: round(v[m], 4) for m in ("precision", "recall", "iou")} for k, v in out.items()})


if __name__ == "__main__":
    main()

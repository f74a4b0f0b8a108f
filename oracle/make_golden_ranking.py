#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- generates tests/golden/ranking_task.json by running the UNMODIFIED reference
``TopKRankingEvaluator`` (/root/reference/src/evaluation/tasks.py:776-901) in the build container on small test
tables with a table-driven recommender (scores with ties, unknown ids -> 0.0, duplicate rows, more items than K).
The reference cannot travel to the GPU box, so its outputs are committed as fixtures; the oracle restatement
(``oracle/pxr_oracle.py::ranking_task_metrics``) and the product (``evaluation.RankingEvaluator``) are pinned to
them in tests/test_ranking_task.py.

    python oracle/make_golden_ranking.py
"""
import contextlib
import io
import json
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pandas as pd

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, "/root/reference")


class TableRecommender:
    def __init__(self, table):
        self.table = table

    def get_item_score(self, user_id, item_id):
        return float(self.table.get(f"{user_id}|{item_id}", 0.0))     # Recommender.get_item_score: 0.0 for unknown ids


def cases():
    rng = np.random.default_rng(20261018)
    out = {}
    # the reference's own known-answer case (tests/unit/src/evaluation/test_tasks.py:113-147)
    out["reference_kat"] = dict(top_k=5, rows=[["u1", "i5"], ["u1", "i2"], ["u1", "i8"]],
                                table={"u1|i5": 0.5, "u1|i2": 0.2, "u1|i8": 0.8})
    # ties (stable over the table order), a duplicate row, an unknown item, an unknown user, n > K
    rows = [["u2", "i3"], ["u1", "i1"], ["u1", "i2"], ["u1", "i3"], ["u2", "i9"], ["u1", "i2"], ["zz", "i1"],
            ["u3", "i4"], ["u2", "i1"], ["u1", "i7"], ["u1", "nope"], ["u2", "i3"], ["u2", "i5"]]
    table = {"u1|i1": 0.25, "u1|i2": 0.75, "u1|i3": 0.25, "u1|i7": 0.0, "u2|i3": 0.5, "u2|i9": 0.5, "u2|i1": 0.625,
             "u2|i5": 0.125, "u3|i4": 0.3}
    out["ties_dups_unknown"] = dict(top_k=3, rows=rows, table=table)
    # random: 40 users, 1..30 items each, quantised scores (many ties), K = 10
    rows, table = [], {}
    for u in range(40):
        n = int(rng.integers(1, 31))
        for it in rng.integers(0, 60, size=n):
            rows.append([f"u{u:03d}", f"i{int(it):03d}"])
            table[f"u{u:03d}|i{int(it):03d}"] = float(rng.integers(0, 16)) / 16.0
    perm = rng.permutation(len(rows))
    out["random_quantised"] = dict(top_k=10, rows=[rows[i] for i in perm], table=table)
    return out


def main():
    from src.evaluation.tasks import TopKRankingEvaluator
    golden = {}
    for name, c in cases().items():
        df = pd.DataFrame(c["rows"], columns=["user_id", "item_id"])
        cfg = SimpleNamespace(recommendation=SimpleNamespace(top_k=c["top_k"]))
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            res = TopKRankingEvaluator(TableRecommender(c["table"]), df, cfg).evaluate()
        res["predictions"] = {u: [[i, float(s)] for i, s in lst] for u, lst in res["predictions"].items()}
        res = {k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in res.items()}
        golden[name] = dict(c, expected=res)
    p = REPO / "tests" / "golden" / "ranking_task.json"
    p.write_text(json.dumps(golden, indent=1))
    print("wrote", p, {k: {m: v["expected"][m] for m in ("avg_avg_rank", "avg_hit_rate_at_k", "avg_ndcg_at_k")} for k, v in golden.items()})


if __name__ == "__main__":
    main()

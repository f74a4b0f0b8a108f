"""Ranking task (``scripts/evaluate.py --eval_task ranking``; reference src/evaluation/tasks.py:776-901).

CPU: the oracle restatement and the product's host logic (``RankingEvaluator`` over a table-driven stand-in for the
recommender, no GPU involved) against outputs of the unmodified reference evaluator (tests/golden/ranking_task.json,
written by oracle/make_golden_ranking.py), including the reference's own known-answer case
(tests/unit/src/evaluation/test_tasks.py:113-147).  GPU: the batched evaluator on the real recommender against the
oracle loop driven by the per-pair reference-style call ``get_item_score``.
"""
import json
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import pxr_oracle as orc
from pixelrec_multimodal_b200.evaluation import RankingEvaluator

GOLDEN = json.loads((Path(__file__).parent / "golden" / "ranking_task.json").read_text())
KEYS = [f"{a}_{m}" for a in ("avg", "std") for m in ("avg_rank", "median_rank", "mrr", "hit_rate_at_k", "ndcg_at_k")]


class TableRecommender:
    """What ``RankingEvaluator`` reads on a recommender: the two id maps and the batched pair scorer."""

    def __init__(self, table):
        users = sorted({k.split("|")[0] for k in table})
        items = sorted({k.split("|")[1] for k in table})
        self.user_index = {u: j for j, u in enumerate(users)}
        self.item_index = {i: j for j, i in enumerate(items)}
        self._users, self._items, self.table = users, items, table

    def score_pairs_batch(self, ui, ii):
        return torch.tensor([self.table.get(f"{self._users[int(u)]}|{self._items[int(i)]}", 0.0) for u, i in zip(ui, ii)],
                            dtype=torch.float32)


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_matches_reference_evaluator(name):
    c = GOLDEN[name]
    df = pd.DataFrame(c["rows"], columns=["user_id", "item_id"])
    groups = df.groupby("user_id")
    users = [str(u) for u, _ in groups]
    items = [[str(i) for i in g["item_id"].tolist()] for _, g in groups]
    res = orc.ranking_task_metrics(users, items, lambda u, i: c["table"].get(f"{u}|{i}", 0.0), c["top_k"])
    for k in KEYS:
        assert res[k] == c["expected"][k], k                                  # same float64 operations: bit-exact
    assert res["num_users_evaluated"] == c["expected"]["num_users_evaluated"]
    assert {u: [[i, s] for i, s in lst] for u, lst in res["predictions"].items()} == c["expected"]["predictions"]


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_product_host_logic_matches_reference_evaluator(name):
    c = GOLDEN[name]
    df = pd.DataFrame(c["rows"], columns=["user_id", "item_id"])
    res = RankingEvaluator(TableRecommender(c["table"]), df, top_k=c["top_k"]).evaluate()
    for k in KEYS:
        assert res[k] == pytest.approx(c["expected"][k], rel=0, abs=1e-12), k
    assert res["num_users_evaluated"] == c["expected"]["num_users_evaluated"]
    assert list(res["predictions"]) == list(c["expected"]["predictions"])      # groupby order of the users
    # the product's scores are fp32 (like the reference model's ``.item()``); the table holds float64 literals
    assert {u: [[i, s] for i, s in lst] for u, lst in res["predictions"].items()} == \
        {u: [[i, float(np.float32(s))] for i, s in lst] for u, lst in c["expected"]["predictions"].items()}


def test_reference_known_answers():
    """tests/unit/src/evaluation/test_tasks.py:130-147, restated."""
    c = GOLDEN["reference_kat"]
    df = pd.DataFrame(c["rows"], columns=["user_id", "item_id"])
    res = RankingEvaluator(TableRecommender(c["table"]), df, top_k=5).evaluate()
    assert res["avg_avg_rank"] == pytest.approx(2.0)
    assert res["avg_median_rank"] == pytest.approx(2.0)
    assert res["avg_mrr"] == pytest.approx(1.0)
    assert res["avg_hit_rate_at_k"] == pytest.approx(1.0)
    assert res["avg_ndcg_at_k"] == pytest.approx(1.0)
    assert res["num_users_evaluated"] == 1
    assert [i for i, _ in res["predictions"]["u1"]] == ["i8", "i5", "i2"]


def test_empty_table_and_workers():
    df = pd.DataFrame({"user_id": [], "item_id": []})
    res = RankingEvaluator(TableRecommender({"u|i": 1.0}), df, top_k=5).evaluate()
    assert res["num_users_evaluated"] == 0 and res["avg_mrr"] == 0.0 and res["predictions"] == {}   # tasks.py:894-897
    with pytest.raises(ValueError):
        RankingEvaluator(TableRecommender({"u|i": 1.0}), df, num_workers=4)


@pytest.mark.gpu
def test_ranking_evaluator_on_gpu_matches_oracle_loop():
    from pixelrec_multimodal_b200 import FastRecommender, ItemFeatureStore, synthetic as syn
    from tests import _cases as cs
    spec = syn.ModelSpec(n_users=60, n_items=400, fusion_type="gated")
    sd, feats = cs.make_workload(spec, syn.SEED + 41)
    model = cs.torch_model_from(spec, sd)
    uids, iids = syn.user_ids(spec.n_users), syn.item_ids(spec.n_items)

    class _DS:
        class _E:
            def __init__(self, c): self.classes_ = np.array(c)
        user_encoder, item_encoder, interactions = _E(uids), _E(iids), None

    store = ItemFeatureStore(torch.from_numpy(feats["tag_idx"]), torch.from_numpy(feats["vis"]), torch.from_numpy(feats["txt"]),
                             torch.from_numpy(feats["num"]))
    rec = FastRecommender(model, _DS(), torch.device("cuda:0"), item_features=store)
    rng = np.random.default_rng(5)
    rows = []
    for u in uids[:24]:
        for it in rng.choice(len(iids), size=int(rng.integers(1, 20)), replace=True):
            rows.append([u, iids[int(it)]])
    rows += [["nobody", iids[0]], [uids[0], "no-such-item"], [uids[1], iids[3]], [uids[1], iids[3]]]
    rows = [rows[i] for i in rng.permutation(len(rows))]
    df = pd.DataFrame(rows, columns=["user_id", "item_id"])
    res = RankingEvaluator(rec, df, top_k=10).evaluate()
    groups = df.groupby("user_id")
    exp = orc.ranking_task_metrics([str(u) for u, _ in groups], [[str(i) for i in g["item_id"].tolist()] for _, g in groups],
                                   rec.get_item_score, 10)
    for k in KEYS:
        assert res[k] == pytest.approx(exp[k], rel=0, abs=1e-12), k
    assert res["num_users_evaluated"] == exp["num_users_evaluated"] == 25
    assert res["predictions"] == exp["predictions"]                            # same kernel, same floats, same stable order
    assert res["predictions"]["nobody"] == [(iids[0], 0.0)]
    # the scores themselves: fp32 kernel vs the fp64 oracle forward (tolerance of the fp32 path, 5e-5)
    cfg = cs.spec_cfg(spec)
    for u in uids[:6]:
        lst = [(i, s) for i, s in res["predictions"][u] if i in set(iids)]
        ci = np.array([iids.index(i) for i, _ in lst])
        sc = orc.forward_pairs(sd, cfg, np.full(len(ci), uids.index(u)), ci, feats["tag_idx"][ci], feats["vis"][ci], feats["txt"][ci], feats["num"][ci])
        assert np.allclose([s for _, s in lst], np.asarray(sc).reshape(-1), atol=5e-5, rtol=0)

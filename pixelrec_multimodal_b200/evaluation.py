"""GPU ranking metrics and the full-catalogue evaluator.

``FullCatalogueEvaluator`` does what the reference's ``use_sampling=False``
docstring promises but its code does not: every user is
scored against the whole catalogue (``get_recommendations(candidates=None)``,
reference ``src/inference/recommender.py:73-79``) and the accuracy block of
``TopKRetrievalEvaluator.evaluate`` (``src/evaluation/tasks.py:567-635``,
``_calculate_ndcg`` :718-747) is evaluated on the resulting top-K lists, on the
GPU (``pxr_metrics``), for several cut-offs at once.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .engine import ranking_metric_sums

_COLS = ("avg_precision_at_k", "avg_recall_at_k", "avg_f1_at_k", "avg_hit_rate_at_k", "avg_ndcg_at_k", "avg_mrr",
         "avg_ndcg_list_ideal_at_k")


def ranking_metrics(topk_idx: torch.Tensor, gt_indptr, gt_idx, ks: Sequence[int]) -> Dict[int, Dict[str, float]]:
    """Means over ALL users of the batch (users without positives contribute
    zeros, tasks.py:589-591, 623-630).  Keys follow tasks.py:623-630;
    ``avg_ndcg_list_ideal_at_k`` is the other NDCG definition the reference ships
    (``src/evaluation/metrics.py:63-100``)."""
    gt_indptr = torch.as_tensor(gt_indptr)
    gt_idx = torch.as_tensor(gt_idx)
    n = int(topk_idx.shape[0])
    sums = ranking_metric_sums(topk_idx, gt_indptr, gt_idx, ks)
    out = {}
    for row, k in zip(sums, sorted(int(k) for k in ks)):
        out[k] = {c: (float(v) / n if n else 0.0) for c, v in zip(_COLS, row)}
        out[k]["num_users_evaluated"] = n
    return out


class FullCatalogueEvaluator:
    """``evaluate()`` returns the result-dict keys of tasks.py:623-635 for
    ``top_k`` (plus ``by_k`` with every requested cut-off)."""

    def __init__(self, recommender, test_data, top_k: int = 50, ks: Optional[Sequence[int]] = None,
                 filter_seen: bool = True, keep_predictions: bool = False, num_workers: int = 0):
        if num_workers and num_workers > 1:
            # the reference forks worker processes holding the model (tasks.py:546-561);
            # a CUDA context cannot be forked
            raise ValueError("num_workers > 1 is not supported on the GPU path (a CUDA context cannot be forked)")
        self.recommender = recommender
        self.top_k = int(top_k)
        self.ks = sorted(set(int(k) for k in (ks or [top_k])) | {self.top_k})
        self.filter_seen = filter_seen
        self.keep_predictions = keep_predictions
        r = recommender
        uu = test_data["user_id"].astype(str).map(r.user_index)
        ii = test_data["item_id"].astype(str).map(r.item_index)
        ok = uu.notna() & ii.notna()
        u = uu[ok].to_numpy(dtype=np.int64)
        i = ii[ok].to_numpy(dtype=np.int64)
        pairs = np.unique(u * (1 << 32) + i)
        u, i = pairs >> 32, pairs & 0xFFFFFFFF
        self.users = np.unique(u)                              # groupby('user_id') order (tasks.py:537)
        remap = {int(x): j for j, x in enumerate(self.users)}
        cnt = np.zeros(len(self.users) + 1, dtype=np.int64)
        np.add.at(cnt, np.array([remap[int(x)] for x in u], dtype=np.int64) + 1, 1)
        self.gt_indptr = np.cumsum(cnt)
        self.gt_idx = i.astype(np.int32)

    def evaluate(self) -> Dict:
        r = self.recommender
        kmax = max(self.ks)
        scores, idx = r.recommend_all(self.users, top_k=kmax, filter_seen=self.filter_seen)
        by_k = ranking_metrics(idx, self.gt_indptr, self.gt_idx, self.ks)
        res = {k: v for k, v in by_k[self.top_k].items() if k != "avg_ndcg_list_ideal_at_k"}
        res["evaluation_method"] = "full_evaluation"
        res["by_k"] = by_k
        if self.keep_predictions:
            s, i = scores.cpu().numpy(), idx.cpu().numpy()
            res["predictions"] = {str(r.user_ids[int(u)]): [(str(r.item_ids[int(b)]), float(a)) for a, b in zip(s[j], i[j]) if b >= 0]
                                  for j, u in enumerate(self.users)}
        return res

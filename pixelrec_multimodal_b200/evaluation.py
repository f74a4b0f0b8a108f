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

from .engine import ranking_metric_sums, sample_candidates, weighted_candidates

_COLS = ("avg_precision_at_k", "avg_recall_at_k", "avg_f1_at_k", "avg_hit_rate_at_k", "avg_ndcg_at_k", "avg_mrr",
         "avg_ndcg_list_ideal_at_k", "avg_precision_hits_over_k", "avg_map_at_k")
# columns of by_k that are not keys of the reference result dict (tasks.py:623-630): the standalone definitions of
# src/evaluation/metrics.py (NDCG normalised by the list's own ideal :63-100, precision = hits / k :35, MAP :102-133)
_EXTRA_COLS = ("avg_ndcg_list_ideal_at_k", "avg_precision_hits_over_k", "avg_map_at_k")


def ranking_metrics(topk_idx: torch.Tensor, gt_indptr, gt_idx, ks: Sequence[int], recall_den=None,
                    n_total: Optional[int] = None) -> Dict[int, Dict[str, float]]:
    """Means over ALL users of the batch (users without positives contribute
    zeros, tasks.py:589-591, 623-630).  Keys follow tasks.py:623-630; the ``_EXTRA_COLS`` are the standalone
    definitions the reference ships in ``src/evaluation/metrics.py``.  ``recall_den``: raw positive-row counts
    (tasks.py:579); ``n_total``: number of users in the mean when it exceeds the rows of ``topk_idx`` (test users
    unknown to the encoder get no list and count as zeros, tasks.py:537-540)."""
    gt_indptr = torch.as_tensor(gt_indptr)
    gt_idx = torch.as_tensor(gt_idx)
    n = int(n_total if n_total is not None else topk_idx.shape[0])
    sums = ranking_metric_sums(topk_idx, gt_indptr, gt_idx, ks,
                               recall_den=None if recall_den is None else torch.as_tensor(recall_den))
    return _by_k(sums, ks, n)


def _by_k(sums, ks, n: int) -> Dict[int, Dict[str, float]]:
    out = {}
    for row, k in zip(sums, sorted(int(k) for k in ks)):
        out[k] = {c: (float(v) / n if n else 0.0) for c, v in zip(_COLS, row)}
        out[k]["num_users_evaluated"] = n
    return out


def build_ground_truth(recommender, test_data):
    """The per-user ground truth of ``TopKRetrievalEvaluator`` (tasks.py:537-540, 322-326, 576-603) as arrays.

    Every distinct ``user_id`` of the test table is one evaluated user (``groupby('user_id')``, string order), known
    to the encoders or not.  For a user the reference keeps ``positive_items`` = ALL its rows as strings: the recall
    denominator is the raw row count (duplicates and items unknown to the item encoder included, :579), hits / NDCG /
    MRR use the SET of those strings (:586-603, ideal DCG over ``min(len(set), K)``).  An unknown item can never be
    recommended, so it only widens the set: it is given a virtual index >= n_items that no list entry equals.

    Returns a dict: ``users`` (encoder indices of the KNOWN test users, string order), ``n_total`` (all distinct test
    users), ``unknown_users`` (their ids), ``gt_indptr / gt_idx`` (the positive SET per known user, ascending, virtual
    indices last), ``recall_den`` (raw row counts), ``pos_indptr / pos_idx`` (known positives only, for samplers)."""
    r = recommender
    uid = test_data["user_id"].astype(str).to_numpy(dtype=object).astype(str)
    iid = test_data["item_id"].astype(str).to_numpy(dtype=object).astype(str)
    users_all, inv = np.unique(uid, return_inverse=True) if len(uid) else (np.zeros(0, dtype=str), np.zeros(0, np.int64))
    umap, imap = r.user_index, r.item_index
    u_enc = np.fromiter((umap.get(str(u), -1) for u in users_all), dtype=np.int64, count=len(users_all))
    known_u = u_enc >= 0
    slot = np.cumsum(known_u) - 1                                  # position among the known users
    rows = known_u[inv] if len(inv) else np.zeros(0, bool)
    seg = slot[inv[rows]] if len(inv) else np.zeros(0, np.int64)
    items = iid[rows]
    n_known = int(known_u.sum())
    n_items = int(r.n_items)
    it = np.fromiter((imap.get(str(i), -1) for i in items), dtype=np.int64, count=len(items))
    unk = it < 0
    if unk.any():
        _, code = np.unique(items[unk], return_inverse=True)
        it[unk] = n_items + code
        if n_items + int(code.max()) >= (1 << 31) - 1:
            raise ValueError("too many unknown test items for int32 virtual indices")
    recall_den = np.bincount(seg, minlength=n_known).astype(np.int32) if n_known else np.zeros(0, np.int32)
    pairs = np.unique(seg * (1 << 32) + it) if len(seg) else np.zeros(0, np.int64)
    pu, pi = pairs >> 32, pairs & 0xFFFFFFFF

    def csr(u, i):
        ptr = np.zeros(n_known + 1, dtype=np.int64)
        if len(u):
            np.add.at(ptr, u + 1, 1)
        return np.cumsum(ptr), i.astype(np.int32)

    gt_indptr, gt_idx = csr(pu, pi)
    kn = pi < n_items
    pos_indptr, pos_idx = csr(pu[kn], pi[kn])
    return dict(users=u_enc[known_u], n_total=len(users_all), unknown_users=[str(u) for u in users_all[~known_u]],
                gt_indptr=gt_indptr, gt_idx=gt_idx, recall_den=recall_den, pos_indptr=pos_indptr, pos_idx=pos_idx)


def novelty_tables(interactions_items: np.ndarray, n_hist_users: int, n_items: int):
    """Per-item float64 tables of ``NoveltyMetrics`` (reference src/evaluation/novelty.py:26-65, 149-206) from the
    interaction table: (self_information, iif, number of items with a popularity entry); NaN = no entry."""
    cnt = np.bincount(np.asarray(interactions_items, dtype=np.int64), minlength=n_items).astype(np.float64)
    total = cnt.sum()
    has = cnt > 0
    si = np.full(n_items, np.nan)
    iif = np.full(n_items, np.nan)
    if total > 0:
        si[has] = -np.log2(np.maximum(cnt[has] / total, 1e-10))
    if n_hist_users > 0:
        iif[has] = np.log(n_hist_users / (cnt[has] + 1e-10))
    return si, iif, int(has.sum())


def beyond_accuracy_metrics(topk_idx: torch.Tensor, si: np.ndarray, iif: np.ndarray, n_pop: int,
                            hist_indptr=None, hist_idx=None) -> Dict[str, float]:
    """Novelty / coverage / personalization over the ranked lists on the GPU (``pxr_novelty_metrics``), keys and
    aggregation of ``TopKRetrievalEvaluator.evaluate`` (tasks.py:674-714).  (``avg_intra_list_similarity`` and the Gini
    coefficient are separate kernels: ``intra_list_similarity`` / ``gini_coefficient`` below.)"""
    import ctypes as C
    from . import _lib
    from .engine import _ptr, _stream
    lib = _lib.load()
    dev = topk_idx.device
    topk_idx = topk_idx.to(torch.int32).contiguous()
    n, k = topk_idx.shape
    n_items = len(si)
    d_si, d_iif = torch.from_numpy(si).to(dev), torch.from_numpy(iif).to(dev)
    ip = torch.as_tensor(hist_indptr).to(device=dev, dtype=torch.int64).contiguous() if hist_indptr is not None else None
    ix = torch.as_tensor(hist_idx).to(device=dev, dtype=torch.int32).contiguous() if hist_idx is not None else None
    out = torch.empty(6, dtype=torch.float64, device=dev)
    nbytes = int(lib.pxr_novelty_bytes(n, n_items))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pxr_novelty_metrics(_ptr(topk_idx), k, n, n_items, _ptr(d_si), _ptr(d_iif), _ptr(ip), _ptr(ix), _ptr(out),
                                     _ptr(ws), nbytes, _stream())
    if rc != 0:
        raise _lib.PxrError(f"pxr_novelty_metrics failed ({rc})")
    s_si, s_iif, s_uniq, s_pn, n_ne, ssq = out.cpu().tolist()
    m = lambda x: float(x / n_ne) if n_ne else 0.0
    if n <= 1:
        pers = 1.0 if n == 1 else 0.0
    else:
        pers = 1.0 - ((ssq - n_ne) / 2.0) / (n * (n - 1) / 2.0)
    return {"avg_self_information": m(s_si), "avg_iif": m(s_iif),
            "avg_catalog_coverage": float(s_uniq / n_pop / n_ne) if (n_ne and n_pop) else 0.0,
            "avg_personalization": float(pers), "avg_personalized_novelty": m(s_pn)}


def gini_coefficient(topk_idx: torch.Tensor, n_items: int, include_zero: bool = False) -> float:
    """Gini coefficient of the per-item recommendation counts of the ranked lists on the GPU (``pxr_gini``):
    ``AdvancedMetrics.calculate_gini_coefficient`` (src/evaluation/advanced_metrics.py:72-105) of
    {item: number of lists holding it}; ``include_zero`` puts the never-recommended items into the distribution."""
    import ctypes as C
    from . import _lib
    from .engine import _ptr, _stream
    lib = _lib.load()
    dev = topk_idx.device
    topk_idx = topk_idx.to(torch.int32).contiguous()
    n, k = topk_idx.shape
    nbytes = int(lib.pxr_gini_bytes(n, int(n_items)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(3, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pxr_gini(_ptr(topk_idx), k, n, int(n_items), int(bool(include_zero)), _ptr(out), _ptr(ws), nbytes, _stream())
    if rc != 0:
        raise _lib.PxrError(f"pxr_gini failed ({rc})")
    return float(out[0].item())


def intra_list_similarity(topk_idx: torch.Tensor, engine=None, embeddings: Optional[torch.Tensor] = None, item_base: int = 0) -> float:
    """Mean over users of the intra-list similarity (``NoveltyMetrics.calculate_diversity``,
    src/evaluation/novelty.py:295-340, averaged as in tasks.py:695-701) on the GPU (``pxr_intra_list_similarity``).
    Embeddings: a caller-provided (n_items, dim) fp32 table, or -- ``engine`` given -- the item records already
    resident for scoring (the projected item-side modality vectors)."""
    import ctypes as C
    from . import _lib
    from .engine import _ptr, _stream
    lib = _lib.load()
    dev = topk_idx.device
    topk_idx = topk_idx.to(torch.int32).contiguous()
    n, k = topk_idx.shape
    if embeddings is not None:
        emb = embeddings.to(device=dev, dtype=torch.float32).contiguous()
        n_rows, dim = int(emb.shape[0]), int(emb.shape[1])
    elif engine is not None:
        emb, n_rows, dim, item_base = None, int(engine.n_rows), 0, 0
    else:
        raise ValueError("intra_list_similarity needs an engine (resident item records) or an embedding table")
    nbytes = int(lib.pxr_ils_bytes(n, n_rows))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pxr_intra_list_similarity(engine._h if engine is not None else None, _ptr(topk_idx), k, n, _ptr(emb), n_rows, dim,
                                           int(item_base), _ptr(out), _ptr(ws), nbytes, _stream())
    if rc != 0:
        raise _lib.PxrError(f"pxr_intra_list_similarity failed ({rc})")
    s, c = out.cpu().tolist()
    return float(s / c) if c else 0.0


# ----------------------------------------------------------------------------- popularity-biased candidate sampling
def sampling_weights(test_items: np.ndarray, n_items: int, strategy: str) -> np.ndarray:
    """Item weights of the 'popularity' / 'popularity_inverse' strategies (src/evaluation/tasks.py:227-243, 266-280):
    rows of the TEST table holding the item (1 for items that never occur there), or the reciprocal."""
    cnt = np.bincount(np.asarray(test_items, dtype=np.int64), minlength=n_items).astype(np.float64)
    cnt[cnt == 0] = 1.0
    return cnt if strategy == "popularity" else 1.0 / cnt


class FullCatalogueEvaluator:
    """``evaluate()`` returns the result-dict keys of tasks.py:623-635 for
    ``top_k`` (plus ``by_k`` with every requested cut-off)."""

    def __init__(self, recommender, test_data, top_k: int = 50, ks: Optional[Sequence[int]] = None,
                 filter_seen: bool = True, keep_predictions: bool = False, num_workers: int = 0,
                 sharded=None, user_block: int = 32768):
        """``sharded``: a ``sharding.ShardedTopK`` when ``recommender`` holds one item-axis shard per rank
        (``torch.distributed`` initialised): the per-shard lists of every user block are exchanged and merged, each
        rank computes the metric sums of its slice of the block, and ONE all-reduce of the (n_ks, 7) float64 sums
        closes the evaluation."""
        self.sharded = sharded
        self.user_block = int(user_block)
        self._metric_sums = ranking_metric_sums
        if num_workers and num_workers > 1:
            # the reference forks worker processes holding the model (tasks.py:546-561);
            # a CUDA context cannot be forked
            raise ValueError("num_workers > 1 is not supported on the GPU path (a CUDA context cannot be forked)")
        self.recommender = recommender
        self.top_k = int(top_k)
        self.ks = sorted(set(int(k) for k in (ks or [top_k])) | {self.top_k})
        self.filter_seen = filter_seen
        self.keep_predictions = keep_predictions
        gt = build_ground_truth(recommender, test_data)
        self.users = gt["users"]                                # known test users, groupby('user_id') order (tasks.py:537)
        self.n_total = gt["n_total"]                            # every distinct test user counts in the means
        self.unknown_users = gt["unknown_users"]
        self.gt_indptr, self.gt_idx, self.recall_den = gt["gt_indptr"], gt["gt_idx"], gt["recall_den"]
        self.pos_indptr, self.pos_idx = gt["pos_indptr"], gt["pos_idx"]

    def novelty(self, topk_idx: torch.Tensor) -> Dict[str, float]:
        """Novelty / coverage / personalization block of ``evaluate`` (tasks.py:637-714) for lists aligned with
        ``self.users``; {} when the recommender has no interaction history (tasks.py:650-652)."""
        r = self.recommender
        if not r.has_history:
            return {}
        ip, ix = r._host_history()
        inter = getattr(r.dataset, "interactions", None)
        if inter is not None and len(inter):
            # popularity counts interaction ROWS (value_counts, tasks.py:646), duplicates included
            ii = inter["item_id"].astype(str).map(r.item_index)
            items_for_counts = ii[ii.notna()].to_numpy(dtype=np.int64)
            n_hist_users = int(inter["user_id"].astype(str).nunique())
        else:
            items_for_counts, n_hist_users = ix, int(np.count_nonzero(np.diff(ip)))
        si, iif, n_pop = novelty_tables(items_for_counts, n_hist_users, r.n_items)
        lens = ip[self.users + 1] - ip[self.users]
        sub_ip = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        sub_ix = np.concatenate([ix[ip[u]:ip[u + 1]] for u in self.users]).astype(np.int32) if len(self.users) and sub_ip[-1] else np.zeros(1, np.int32)
        out = beyond_accuracy_metrics(topk_idx[:, :self.top_k], si, iif, n_pop, sub_ip, sub_ix)
        lists = topk_idx[:, :self.top_k].contiguous()
        out["gini_coefficient"] = gini_coefficient(lists, r.n_items)                    # advanced_metrics.py:72-105
        if r.item_lo == 0 and r.item_hi == r.n_items:                                  # needs every record on this rank
            out["avg_intra_list_similarity"] = intra_list_similarity(lists, engine=r.engine())   # novelty.py:295-340, tasks.py:695-701
        return out

    def _evaluate_sharded(self, kmax: int, want_lists: bool):
        """Item-sharded ranks in lock step: every rank ends up with the final lists of the users it owns in each block
        (``ShardedTopK.recommend_blocks_owned``), computes their metric sums, and ONE all-reduce of the (n_ks, 9) float64
        sums closes the evaluation.  ``want_lists``: also gather the lists (predictions / novelty)."""
        import torch.distributed as dist
        from .sharding import gather_owned, owned_slice
        world, rank = dist.get_world_size(self.sharded.group), dist.get_rank(self.sharded.group)
        n = len(self.users)
        blocks = [self.users[i:i + self.user_block] for i in range(0, n, self.user_block)]
        sums = np.zeros((len(self.ks), len(_COLS)))
        outs_s, outs_i, row = [], [], 0
        for (s, i), blk in zip(self.sharded.recommend_blocks_owned(blocks, kmax, self.filter_seen), blocks):
            lo, hi = owned_slice(len(blk), world, rank)
            if hi > lo:
                ip = self.gt_indptr[row + lo:row + hi + 1]
                sums += self._metric_sums(i, torch.from_numpy(ip - ip[0]),
                                          torch.from_numpy(self.gt_idx[ip[0]:ip[-1]] if ip[-1] > ip[0] else np.zeros(1, np.int32)), self.ks,
                                          recall_den=torch.from_numpy(self.recall_den[row + lo:row + hi]))
            if want_lists:
                gs, gi = gather_owned(s, i, len(blk), self.sharded.group)
                outs_s.append(gs); outs_i.append(gi)
            row += len(blk)
        dev = self.recommender.device
        t = torch.from_numpy(sums).to("cpu" if dist.get_backend(self.sharded.group) == "gloo" else dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.sharded.group)
        by_k = _by_k(t.cpu().numpy(), self.ks, self.n_total)
        return (torch.cat(outs_s) if outs_s else None), (torch.cat(outs_i) if outs_i else None), by_k

    def evaluate(self, novelty: bool = False) -> Dict:
        r = self.recommender
        kmax = max(self.ks)
        if self.sharded is not None:
            scores, idx, by_k = self._evaluate_sharded(kmax, want_lists=novelty or self.keep_predictions)
        else:
            scores, idx = r.recommend_all(self.users, top_k=kmax, filter_seen=self.filter_seen)
            by_k = ranking_metrics(idx, self.gt_indptr, self.gt_idx, self.ks, recall_den=self.recall_den, n_total=self.n_total)
        res = {k: v for k, v in by_k[self.top_k].items() if k not in _EXTRA_COLS}
        res["evaluation_method"] = "full_evaluation"
        res["by_k"] = by_k
        if novelty:
            res.update(self.novelty(idx))
        if self.keep_predictions:
            s, i = scores.cpu().numpy(), idx.cpu().numpy()
            res["predictions"] = {str(r.user_ids[int(u)]): [(str(r.item_ids[int(b)]), float(a)) for a, b in zip(s[j], i[j]) if b >= 0]
                                  for j, u in enumerate(self.users)}
            res["predictions"].update({u: [] for u in self.unknown_users})      # get_recommendations -> [] (recommender.py:65-67)
        return res


class SampledRetrievalEvaluator(FullCatalogueEvaluator):
    """The reference's default protocol (``scripts/evaluate.py --use_sampling``,
    ``TopKRetrievalEvaluator`` with ``sampling_strategy='random'``, ``src/evaluation/tasks.py:181-224,
    310-364``) on the GPU: per user the test positives plus ``num_negatives`` uniformly sampled non-positive
    items form a shuffled candidate list (``pxr_sample_candidates``), the recommender ranks exactly those
    (``filter_seen=False``), and the accuracy block of ``evaluate`` (:567-635) is computed on the top-K lists.
    Sampling is a pure function of ``(seed, user index)``; the reference seeds with Python's salted ``hash()`` and is
    not reproducible across processes.  ``sampling_strategy`` 'popularity' / 'popularity_inverse' (tasks.py:225-308) draw the
    negatives with probability proportional to the test-set item count / its reciprocal (``weighted_candidates``)."""

    def __init__(self, recommender, test_data, top_k: int = 50, ks: Optional[Sequence[int]] = None,
                 num_negatives: int = 100, sampling_strategy: str = "random", seed: int = 20261018,
                 keep_predictions: bool = False, num_workers: int = 0, user_block: int = 65536):
        if sampling_strategy not in ("random", "popularity", "popularity_inverse"):
            raise ValueError("sampling_strategy must be 'random', 'popularity' or 'popularity_inverse' (tasks.py:221-308)")
        super().__init__(recommender, test_data, top_k=top_k, ks=ks, filter_seen=False,
                         keep_predictions=keep_predictions, num_workers=num_workers)
        self.num_negatives = int(num_negatives)
        self.seed = int(seed)
        self.user_block = int(user_block)
        self.sampling_strategy = sampling_strategy
        self._weights = None
        if sampling_strategy != "random":
            # weights come from the item counts of the TEST table (tasks.py:227, 266), items never seen there count 1
            ii = test_data["item_id"].astype(str).map(recommender.item_index)
            self._weights = sampling_weights(ii[ii.notna()].to_numpy(dtype=np.int64), recommender.n_items, sampling_strategy)

    def candidates(self, lo: int = 0, hi: Optional[int] = None):
        """(n, stride) int32 candidate item indices (device) of evaluated users [lo, hi)."""
        r = self.recommender
        hi = len(self.users) if hi is None else hi
        dev = r.device
        indptr = torch.from_numpy(self.pos_indptr[lo:hi + 1] - self.pos_indptr[lo]).to(dev)
        idx = torch.from_numpy(self.pos_idx[self.pos_indptr[lo]:self.pos_indptr[hi]]).to(dev)
        if idx.numel() == 0:
            idx = torch.zeros(1, dtype=torch.int32, device=dev)
        max_pos = int(np.diff(self.pos_indptr).max()) if len(self.users) else 0
        stride = max(1, min(1024, max_pos + self.num_negatives))
        if self._weights is not None:
            return weighted_candidates(torch.from_numpy(self.users[lo:hi]).to(dev), indptr, idx,
                                       torch.from_numpy(self._weights).to(dev), self.num_negatives, self.seed, stride)
        return sample_candidates(torch.from_numpy(self.users[lo:hi]).to(dev), indptr, idx, r.n_items, self.num_negatives,
                                 self.seed, stride)

    def evaluate(self) -> Dict:
        r = self.recommender
        kmax = max(self.ks)
        sums = np.zeros((len(self.ks), len(_COLS)))
        preds = {}
        for lo in range(0, len(self.users), self.user_block):
            hi = min(len(self.users), lo + self.user_block)
            cand, _ = self.candidates(lo, hi)
            s, items = r.rank_candidates(self.users[lo:hi], cand, top_k=kmax)
            g_idx = self.gt_idx[self.gt_indptr[lo]:self.gt_indptr[hi]]
            sums += ranking_metric_sums(items, torch.from_numpy(self.gt_indptr[lo:hi + 1] - self.gt_indptr[lo]),
                                        torch.from_numpy(g_idx if len(g_idx) else np.zeros(1, np.int32)), self.ks,
                                        recall_den=torch.from_numpy(self.recall_den[lo:hi]))
            if self.keep_predictions:
                hs, hi_ = s.cpu().numpy(), items.cpu().numpy()
                for j, u in enumerate(self.users[lo:hi]):
                    preds[str(r.user_ids[int(u)])] = [(str(r.item_ids[int(b)]), float(a)) for a, b in zip(hs[j], hi_[j]) if b >= 0]
        by_k = _by_k(sums, self.ks, self.n_total)
        res = {k: v for k, v in by_k[self.top_k].items() if k not in _EXTRA_COLS}
        res["evaluation_method"] = "negative_sampling"
        res["by_k"] = by_k
        if self.keep_predictions:
            preds.update({u: [] for u in self.unknown_users})
            res["predictions"] = preds
        return res


class RankingEvaluator:
    """The reference's ranking task (``scripts/evaluate.py --eval_task ranking``, ``TopKRankingEvaluator.evaluate``,
    ``src/evaluation/tasks.py:776-901``) on the GPU path: every (user, test item) pair of the table is scored in one
    batched ``pxr_score_pairs`` launch instead of one ``get_item_score`` forward per pair (:812-820); unknown users
    or items score 0.0 like the reference call (``recommender.py:119-125``).  Inside a user the items are ordered by
    score, descending and stable over the table order (Python's ``sort(reverse=True)``, :830), and every test item
    is relevant, so the per-user values follow :835-854 -- ranks 1..n, MRR 1/ranks[0], hit rate = ranks <= K over
    n, NDCG of the sorted list against the SET of test items.  Result keys, means and population standard
    deviations as :877-897; ``predictions`` holds the sorted (item, score) lists (the reference stores the list
    object it then sorts in place, :827-830).  Users appear in ``groupby('user_id')`` order of the string ids."""

    def __init__(self, recommender, test_data, top_k: int = 50, keep_predictions: bool = True, num_workers: int = 0):
        if num_workers and num_workers > 1:
            raise ValueError("num_workers > 1 is not supported on the GPU path (a CUDA context cannot be forked)")
        self.recommender = recommender
        self.test_data = test_data
        self.top_k = int(top_k)
        self.keep_predictions = keep_predictions

    def evaluate(self) -> Dict:
        r = self.recommender
        uid = self.test_data["user_id"].astype(str).to_numpy(dtype=object)
        iid = self.test_data["item_id"].astype(str).to_numpy(dtype=object)
        users, inv = np.unique(uid.astype(str), return_inverse=True)
        order = np.argsort(inv, kind="stable")                       # rows grouped by user, table order inside a user
        seg, items = inv[order], iid[order]
        n_rows, n_users = len(items), len(users)
        umap, imap = r.user_index, r.item_index
        u_of_user = np.fromiter((umap.get(str(u), -1) for u in users), dtype=np.int64, count=n_users)
        ui = u_of_user[seg] if n_rows else np.zeros(0, np.int64)
        ii = np.fromiter((imap.get(str(i), -1) for i in items), dtype=np.int64, count=n_rows)
        known = (ui >= 0) & (ii >= 0)
        scores = np.zeros(n_rows, dtype=np.float32)                  # get_item_score -> 0.0 for unknown ids
        if known.any():
            scores[known] = r.score_pairs_batch(ui[known], ii[known]).cpu().numpy()
        perm = np.lexsort((-scores, seg))                            # stable: ties keep the table order
        cnt = np.bincount(seg, minlength=n_users).astype(np.int64) if n_rows else np.zeros(n_users, np.int64)
        pair_codes = np.unique(np.stack([seg, np.unique(items.astype(str), return_inverse=True)[1]], 1), axis=0) if n_rows else np.zeros((0, 2), np.int64)
        n_set = np.bincount(pair_codes[:, 0], minlength=n_users).astype(np.int64)   # len(set(test_items))
        k = max(self.top_k, 0)
        gains = np.concatenate([[0.0], np.cumsum(1.0 / np.log2(np.arange(1, k + 1) + 1.0))])   # left-to-right sums of :737-744
        nf = cnt.astype(np.float64)
        ideal = gains[np.minimum(n_set, k)]
        per = {
            "avg_rank": (nf + 1.0) / 2.0,                            # np.mean / np.median of 1..n
            "median_rank": (nf + 1.0) / 2.0,
            "mrr": np.ones(n_users),
            "hit_rate_at_k": np.minimum(cnt, k) / nf,
            "ndcg_at_k": np.divide(gains[np.minimum(cnt, k)], ideal, out=np.zeros(n_users), where=ideal > 0),
        }
        res: Dict = {}
        for name, vals in per.items():
            res[f"avg_{name}"] = float(np.mean(vals)) if n_users else 0.0
            res[f"std_{name}"] = float(np.std(vals)) if n_users else 0.0
        res["num_users_evaluated"] = int(n_users)
        if self.keep_predictions:
            s_sorted, i_sorted = scores[perm], items[perm]
            bounds = np.concatenate([[0], np.cumsum(cnt)])
            res["predictions"] = {str(users[j]): [(str(i_sorted[t]), float(s_sorted[t])) for t in range(bounds[j], bounds[j + 1])]
                                  for j in range(n_users)}
        return res

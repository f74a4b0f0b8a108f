"""FastRecommender: drop-in for the reference inference ``Recommender``
(reference ``src/inference/recommender.py:20-293``) on the full-catalogue path.

Same constructor, same ``get_recommendations`` / ``get_item_score`` /
``print_cache_stats`` / ``clear_cache`` surface and the same duck-typed
``.dataset`` / ``.model`` attributes the evaluators read
(``src/evaluation/tasks.py:171-175, 349-354, 816``), plus the batched
``recommend_all`` the per-user string API can never match for throughput.

What changes underneath: item features are uploaded and projected ONCE
(``pxr_precompute_items``) instead of being re-stacked and re-uploaded for every
user (recommender.py:162-191); string<->index mapping uses dictionaries built
once instead of per-call ``LabelEncoder.transform`` scans (recommender.py:64,
76, 191); user histories become one CSR; scoring + top-K run in libpxr.so.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import PxrEngine, check_index_range, merge_topk
from .model import FastMultimodalRecommender


@dataclass
class ItemFeatureStore:
    """Dense per-item cached features in item-encoder order (row i = item index i):
    the post-backbone embeddings the hoisted frozen backbones emit."""

    tag_idx: torch.Tensor                       # (NI,) int64
    vis: Optional[torch.Tensor] = None          # (NI, Dv) fp32
    txt: Optional[torch.Tensor] = None          # (NI, Dl) fp32
    num: Optional[torch.Tensor] = None          # (NI, F)  fp32
    missing: Optional[np.ndarray] = None        # (NI,) bool: items whose features could not be fetched

    @property
    def n_items(self) -> int:
        return int(self.tag_idx.shape[0])

    @staticmethod
    def from_dataset(dataset, item_ids: Sequence[str], model: FastMultimodalRecommender) -> "ItemFeatureStore":
        """Pull one feature dict per item the way the reference does
        (recommender.py:239-269: dataset.feature_cache first, then
        dataset._get_item_features) and stack them once.  Expected keys are those
        of reference dataset.py:264-303 with cached features in place of raw
        inputs: 'image' (Dv,), 'text_input_ids' (Dl,) float, 'numerical_features'
        (F,), 'tag_idx' ()."""
        n = len(item_ids)
        tag = torch.zeros(n, dtype=torch.int64)
        vis = torch.zeros(n, model.vision_dim) if model.vision_dim else None
        txt = torch.zeros(n, model.language_dim) if model.language_dim else None
        num = torch.zeros(n, model.num_numerical_features) if model.num_numerical_features else None
        missing = np.zeros(n, dtype=bool)
        cache = getattr(dataset, "feature_cache", None)
        for i, iid in enumerate(item_ids):
            feats = None
            if cache is not None:
                try:
                    feats = cache.get(iid)
                except Exception:
                    feats = None
            if not feats and hasattr(dataset, "_get_item_features"):
                try:
                    feats = dataset._get_item_features(iid)
                except Exception:
                    feats = None
            if not feats:
                missing[i] = True
                continue
            tag[i] = int(feats["tag_idx"]) if "tag_idx" in feats else 0
            if vis is not None:
                vis[i] = torch.as_tensor(feats["image"], dtype=torch.float32).reshape(-1)
            if txt is not None:
                txt[i] = torch.as_tensor(feats["text_input_ids"], dtype=torch.float32).reshape(-1)
            if num is not None:
                num[i] = torch.as_tensor(feats["numerical_features"], dtype=torch.float32).reshape(-1)
        return ItemFeatureStore(tag, vis, txt, num, missing if missing.any() else None)


def build_history_csr(user_index: Dict[str, int], item_index: Dict[str, int], interactions,
                      n_users: int) -> Tuple[np.ndarray, np.ndarray]:
    """One CSR of train histories (indptr int64, item idx int32 ascending per
    user) replacing the per-call pandas mask of reference dataset.py:462-476."""
    uu = interactions["user_id"].astype(str).map(user_index)
    ii = interactions["item_id"].astype(str).map(item_index)
    ok = uu.notna() & ii.notna()
    u = uu[ok].to_numpy(dtype=np.int64)
    i = ii[ok].to_numpy(dtype=np.int64)
    pairs = np.unique(u * (1 << 32) + i)
    u, i = pairs >> 32, pairs & 0xFFFFFFFF
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.add.at(indptr, u + 1, 1)
    np.cumsum(indptr, out=indptr)
    return indptr, i.astype(np.int32)


class FastRecommender:
    def __init__(self, model: FastMultimodalRecommender, dataset, device: torch.device,
                 cache_max_items: int = 1000, cache_dir: Optional[str] = None, cache_to_disk: bool = False,
                 item_features: Optional[ItemFeatureStore] = None,
                 history: Optional[Tuple[np.ndarray, np.ndarray]] = None,
                 item_range: Optional[Tuple[int, int]] = None, user_block: int = 8192,
                 n_users: Optional[int] = None, n_items: Optional[int] = None):
        self.model = model
        self.dataset = dataset
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FastRecommender needs a CUDA device: the scoring path has no CPU fallback")
        self.model.to(self.device)
        self.model.eval()
        self.cache_max_items = cache_max_items
        self.feature_cache: Dict[str, Dict[str, torch.Tensor]] = {}
        logging.basicConfig(level=logging.INFO)
        self.logger = logging.getLogger(__name__)
        self.user_block = int(user_block)

        ucls = getattr(getattr(dataset, "user_encoder", None), "classes_", None)
        icls = getattr(getattr(dataset, "item_encoder", None), "classes_", None)
        # id tables are kept as given (no per-element copies: 8.9 M users at Pixel8M scale); the
        # string -> index dictionaries are built on first use by the string API only
        self.user_ids = ucls if ucls is not None else []
        self.item_ids = icls if icls is not None else []
        self.n_users = int(n_users if n_users is not None else len(self.user_ids))
        self.n_items = int(n_items if n_items is not None else len(self.item_ids))
        self._user_index: Optional[Dict[str, int]] = None
        self._item_index: Optional[Dict[str, int]] = None

        self.items = item_features if item_features is not None else \
            ItemFeatureStore.from_dataset(dataset, [str(c) for c in self.item_ids], model)
        if self.items.n_items != self.n_items:
            raise ValueError(f"item feature store has {self.items.n_items} rows, item encoder has {self.n_items}")
        if self.items.missing is not None:
            # reference: items whose features cannot be fetched are not sent through the model and score 0.0
            # (recommender.py:199-201, 229-230); here they carry a per-item flag the kernels honour
            self.logger.warning(f"{int(np.asarray(self.items.missing).sum())} items have no cached features; they score 0.0")

        # train histories -> one CSR resident on the device (global item indices, ascending per user)
        if history is None and getattr(dataset, "interactions", None) is not None and self.n_users:
            history = build_history_csr(self.user_index, self.item_index, dataset.interactions, self.n_users)
        self._h_hist = None                                     # host copy, made on demand
        if history is not None:
            ip, ix = history
            self._d_hist_indptr = torch.as_tensor(ip).to(device=self.device, dtype=torch.int64).contiguous()
            self._d_hist_idx = torch.as_tensor(ix).to(device=self.device, dtype=torch.int32).contiguous()
            if self._d_hist_idx.numel() == 0:
                self._d_hist_idx = torch.zeros(1, dtype=torch.int32, device=self.device)
        else:
            self._d_hist_indptr, self._d_hist_idx = None, None

        lo, hi = item_range if item_range is not None else (0, self.n_items)
        self.item_lo, self.item_hi = int(lo), int(hi)
        self._engine: Optional[PxrEngine] = None
        self._engine_token = None

    @property
    def user_index(self) -> Dict[str, int]:
        if self._user_index is None:
            self._user_index = {str(u): i for i, u in enumerate(self.user_ids)}
        return self._user_index

    @property
    def item_index(self) -> Dict[str, int]:
        if self._item_index is None:
            self._item_index = {str(it): i for i, it in enumerate(self.item_ids)}
        return self._item_index

    @property
    def has_history(self) -> bool:
        return self._d_hist_indptr is not None

    def device_history(self):
        """(indptr int64, idx int32) of the resident train-history CSR, or (None, None)."""
        return self._d_hist_indptr, self._d_hist_idx

    def _host_history(self):
        if self._h_hist is None:
            self._h_hist = (self._d_hist_indptr.cpu().numpy(), self._d_hist_idx.cpu().numpy())
        return self._h_hist

    # -------------------------------------------------------------- catalogue
    def engine(self) -> PxrEngine:
        """Engine with this recommender's item range precomputed (K1+K2 run once)."""
        eng = self.model.engine("catalogue")
        # the records live in the ENGINE (one per model): the token is kept there, so a second recommender on the same
        # model with another item range re-runs the precompute instead of scoring against the other one's shard
        token = (id(self), self.model._engines["catalogue"][1], self.item_lo, self.item_hi)
        if eng.items_token != token:
            lo, hi = self.item_lo, self.item_hi
            sl = slice(lo, hi)
            eng.precompute_items(self.model.item_embedding.weight.detach(), self.items.tag_idx[sl],
                                 None if self.items.vis is None else self.items.vis[sl],
                                 None if self.items.txt is None else self.items.txt[sl],
                                 None if self.items.num is None else self.items.num[sl],
                                 item_idx=None, item_base=lo, n_rows=hi - lo)
            if self.items.missing is not None:
                eng.set_missing_items(torch.from_numpy(np.ascontiguousarray(np.asarray(self.items.missing)[sl]).astype(np.uint8)))
            eng.items_token = token
            self._engine = eng
        return eng

    def rescore_engine(self) -> PxrEngine:
        """Engine holding the fp32 records of the WHOLE catalogue (records only: 1 280 B per item), used by an item-sharded
        rank to re-score the merged candidate lists of the users it owns (exact mode, ``ShardedTopK``)."""
        eng = self.model.engine("rescore")
        token = ("rescore", self.model._engines["rescore"][1], 0, self.n_items)
        if eng.items_token != token:
            eng.set_records_only(True)
            eng.precompute_items(self.model.item_embedding.weight.detach(), self.items.tag_idx, self.items.vis, self.items.txt,
                                 self.items.num, item_idx=None, item_base=0, n_rows=self.n_items)
            if self.items.missing is not None:
                eng.set_missing_items(torch.from_numpy(np.ascontiguousarray(np.asarray(self.items.missing)).astype(np.uint8)))
            eng.items_token = token
        return eng

    @torch.no_grad()
    def rescore(self, user_indices, cand_idx: torch.Tensor, top_k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(n, 64 * pages) candidate lists of global item indices (-1 padded) -> exact (fp32) top-K, ties -> lower index."""
        users = torch.as_tensor(np.asarray(user_indices.cpu() if isinstance(user_indices, torch.Tensor) else user_indices, dtype=np.int64)).to(self.device)
        if users.numel() == 0:
            return (torch.empty((0, top_k), dtype=torch.float32, device=self.device),
                    torch.empty((0, top_k), dtype=torch.int32, device=self.device))
        return self.rescore_engine().rescore_topk(self.model.user_embedding.weight.detach(), users, cand_idx, top_k)

    # ------------------------------------------------------------ batched API
    def _seen_csr_for(self, users: np.ndarray, d_users: torch.Tensor):
        """(seen_indptr, seen_idx) device tensors for ``pxr_score_topk``: offsets
        of each user's ascending item list inside ``seen_idx`` (global indices)."""
        if self._d_hist_indptr is None:
            return torch.zeros(len(users) + 1, dtype=torch.int64, device=self.device), \
                torch.zeros(1, dtype=torch.int32, device=self.device)
        if len(users) and (len(users) == 1 or bool(np.all(np.diff(users) == 1))):
            # contiguous user block: views of the resident CSR, nothing is copied
            u0 = int(users[0])
            return self._d_hist_indptr[u0:u0 + len(users) + 1], self._d_hist_idx
        starts = self._d_hist_indptr[d_users]
        lens = self._d_hist_indptr[d_users + 1] - starts
        indptr = torch.zeros(len(users) + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(lens, 0, out=indptr[1:])
        total = int(indptr[-1])
        if total == 0:
            return indptr, torch.zeros(1, dtype=torch.int32, device=self.device)
        pos = torch.arange(total, device=self.device) + torch.repeat_interleave(starts - indptr[:-1], lens)
        return indptr, self._d_hist_idx[pos]

    @torch.no_grad()
    def recommend_all(self, user_indices, top_k: int = 10, filter_seen: bool = True, raw: bool = False
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Every user of ``user_indices`` (encoder indices) against every item of
        this recommender's item range.  Returns device tensors
        (scores (n, K) fp32 descending, item indices (n, K) int32, -inf / -1 padded).
        ``raw``: the fused kernel's 16-bit lists as they are (no fp32 re-score), e.g. the 64-slot per-shard lists an
        item-sharded job exchanges before the owning rank re-scores the merged candidates."""
        eng = self.engine()
        if raw and eng.rescore:
            eng.set_rescore(False)
            try:
                return self.recommend_all(user_indices, top_k, filter_seen, raw=False)
            finally:
                eng.set_rescore(True)
        users = np.asarray(user_indices.cpu() if isinstance(user_indices, torch.Tensor) else user_indices,
                           dtype=np.int64)
        uemb = self.model.user_embedding.weight.detach()
        if len(users) and (int(users.min()) < 0 or int(users.max()) >= int(uemb.shape[0])):
            raise IndexError(f"user index out of range: [{int(users.min())}, {int(users.max())}] not inside [0, {int(uemb.shape[0])})")
        outs_s, outs_i = [], []
        for u0 in range(0, len(users), self.user_block):
            blk = users[u0:u0 + self.user_block]
            d_users = torch.from_numpy(np.ascontiguousarray(blk)).to(self.device, non_blocking=True)
            if filter_seen:
                indptr, idx = self._seen_csr_for(blk, d_users)
                s, i = eng.score_topk(uemb, d_users, top_k, indptr, idx)
            else:
                s, i = eng.score_topk(uemb, d_users, top_k)
            outs_s.append(s)
            outs_i.append(i)
        if not outs_s:
            return (torch.empty((0, top_k), dtype=torch.float32, device=self.device),
                    torch.empty((0, top_k), dtype=torch.int32, device=self.device))
        return torch.cat(outs_s), torch.cat(outs_i)

    @torch.no_grad()
    def rank_candidates(self, user_indices, candidates: torch.Tensor, top_k: int = 10
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Batched ``get_recommendations(candidates=..., filter_seen=False)`` (reference
        src/inference/recommender.py:81-106): row r of ``candidates`` ((n, C) int32 item indices, -1 padded) is
        ranked for user ``user_indices[r]`` -- stable descending order over the candidate order, first ``top_k``.
        Returns device tensors (scores (n, K), item indices (n, K) int32, -inf / -1 padded)."""
        eng = self.engine()
        if self.item_lo != 0 or self.item_hi != self.n_items:
            raise ValueError("rank_candidates needs the whole catalogue on this recommender (no item-axis shard)")
        users = torch.as_tensor(user_indices, dtype=torch.int64, device=self.device)
        cand = candidates.to(device=self.device, dtype=torch.int32)
        check_index_range(users, int(self.model.user_embedding.weight.shape[0]), "user index")
        if cand.numel() and int(cand.max()) >= self.n_items:
            raise IndexError(f"candidate item index {int(cand.max())} outside the catalogue ({self.n_items} items)")
        n, C_ = cand.shape
        valid = cand >= 0
        rows, cols = valid.nonzero(as_tuple=True)
        scores = torch.full((n, C_), float("-inf"), dtype=torch.float32, device=self.device)
        if rows.numel():
            sc = eng.score_pairs(self.model.user_embedding.weight.detach(), users[rows], cand[rows, cols].to(torch.int64))
            scores[rows, cols] = sc
        s, pos = eng.topk_rows(scores, top_k)
        items = torch.where(pos >= 0, torch.gather(cand, 1, pos.clamp_min(0).to(torch.int64)), torch.full_like(pos, -1))
        return s, items

    def score_pairs_batch(self, user_indices, item_indices) -> torch.Tensor:
        """Batched ``get_item_score`` (reference src/inference/recommender.py:112-141, one forward per (user, item)):
        fp32 device tensor with the score of every (user index, global item index) pair, in ONE ``pxr_score_pairs``
        launch.  Used by the ranking task (``evaluation.RankingEvaluator``)."""
        eng = self.engine()
        users = torch.as_tensor(user_indices, dtype=torch.int64, device=self.device)
        items = torch.as_tensor(item_indices, dtype=torch.int64, device=self.device)
        if users.shape != items.shape or users.dim() != 1:
            raise ValueError("user_indices and item_indices must be 1-D and of equal length")
        if items.numel() == 0:
            return torch.empty(0, dtype=torch.float32, device=self.device)
        if int(items.min()) < self.item_lo or int(items.max()) >= self.item_hi:
            raise ValueError("items outside this recommender's item range")
        check_index_range(users, int(self.model.user_embedding.weight.shape[0]), "user index")
        return eng.score_pairs(self.model.user_embedding.weight.detach(), users, items - self.item_lo)

    # --------------------------------------------------------- reference API
    def get_recommendations(self, user_id: str, top_k: int = 10, filter_seen: bool = True,
                            candidates: Optional[List[str]] = None) -> List[Tuple[str, float]]:
        """reference recommender.py:52-110 — same semantics: unknown user / no
        candidates -> []; candidate order is encoder order (or the given list
        filtered to known items); seen items dropped; STABLE descending sort, so
        ties keep candidate order; first ``top_k``; Python floats."""
        user_id = str(user_id)
        if not len(self.user_ids) or not len(self.item_ids):
            self.logger.warning("User / item encoder not properly initialized.")
            return []
        u = self.user_index.get(user_id)
        if u is None:
            self.logger.warning(f"User '{user_id}' not found in the trained user encoder.")
            return []
        if candidates is None:
            if filter_seen and not self.has_history:
                seen = self._get_user_interactions(user_id)
                return self._recommend_with_seen_set(u, top_k, seen)
            s, i = self.recommend_all(np.array([u]), top_k=top_k, filter_seen=filter_seen)
            s, i = s[0].cpu().numpy(), i[0].cpu().numpy()
            return [(str(self.item_ids[int(ii)]), float(ss)) for ss, ii in zip(s, i) if ii >= 0]
        cand = [str(c) for c in candidates if str(c) in self.item_index]
        if not cand:
            self.logger.info(f"No valid candidate items found for user '{user_id}'.")
            return []
        if filter_seen:
            seen = self._get_user_interactions(user_id)
            cand = [c for c in cand if c not in seen]
        if not cand:
            self.logger.info(f"All candidate items for user '{user_id}' have been filtered.")
            return []
        scores = self._score_items_batch(u, cand)
        pairs = list(zip(cand, scores))
        pairs.sort(key=lambda x: x[1], reverse=True)       # stable, as recommender.py:105
        return pairs[:top_k]

    def _recommend_with_seen_set(self, u: int, top_k: int, seen: Iterable[str]):
        idx = np.array(sorted({self.item_index[s] for s in seen if s in self.item_index}), dtype=np.int32)
        eng = self.engine()
        d_users = torch.tensor([u], dtype=torch.int64, device=self.device)
        indptr = torch.tensor([0, len(idx)], dtype=torch.int64, device=self.device)
        d_idx = torch.from_numpy(idx).to(self.device) if len(idx) else torch.zeros(1, dtype=torch.int32, device=self.device)
        s, i = eng.score_topk(self.model.user_embedding.weight.detach(), d_users, top_k, indptr, d_idx)
        s, i = s[0].cpu().numpy(), i[0].cpu().numpy()
        return [(str(self.item_ids[int(ii)]), float(ss)) for ss, ii in zip(s, i) if ii >= 0]

    def get_item_score(self, user_id: str, item_id: str) -> float:
        """reference recommender.py:112-141: 0.0 for unknown ids."""
        u = self.user_index.get(str(user_id))
        if u is None or str(item_id) not in self.item_index:
            return 0.0
        return self._score_items_batch(u, [str(item_id)])[0]

    def _score_items_batch(self, u: int, item_ids_str: List[str]) -> List[float]:
        """reference recommender.py:144-236 for one user and explicit items."""
        if not item_ids_str:
            return []
        eng = self.engine()
        gi = np.array([self.item_index[i] for i in item_ids_str], dtype=np.int64)
        inside = (gi >= self.item_lo) & (gi < self.item_hi)
        if not inside.all():
            raise ValueError("candidate items outside this recommender's item range")
        rows = torch.from_numpy(gi - self.item_lo).to(self.device)
        users = torch.full((len(gi),), u, dtype=torch.int64, device=self.device)
        out = eng.score_pairs(self.model.user_embedding.weight.detach(), users, rows)
        return [float(x) for x in out.cpu().tolist()]

    def _get_item_features(self, item_id_str: str):
        """reference recommender.py:239-269 (evaluators call this, tasks.py:455)."""
        i = self.item_index.get(str(item_id_str))
        if i is None or (self.items.missing is not None and bool(self.items.missing[i])):
            return None
        feats = {"tag_idx": self.items.tag_idx[i]}
        if self.items.vis is not None:
            feats["image"] = self.items.vis[i]
        if self.items.txt is not None:
            feats["text_input_ids"] = self.items.txt[i]
            feats["text_attention_mask"] = torch.ones(1, dtype=torch.long)
        if self.items.num is not None:
            feats["numerical_features"] = self.items.num[i]
        return feats

    def _get_user_interactions(self, user_id_str: str) -> set:
        """reference recommender.py:271-281."""
        if self.has_history:
            u = self.user_index.get(str(user_id_str))
            if u is None:
                return set()
            ip, ix = self._host_history()
            return {str(self.item_ids[int(j)]) for j in ix[ip[u]:ip[u + 1]]}
        try:
            return set(self.dataset.get_user_history(str(user_id_str)))
        except Exception as e:  # same swallow-and-log as the reference
            self.logger.error(f"Error getting user interactions for user '{user_id_str}': {e}")
            return set()

    def print_cache_stats(self):
        print(f"Feature store: {self.items.n_items} items resident on {self.device}")
        print(f"Cache capacity: {self.cache_max_items}")

    def clear_cache(self):
        self.feature_cache.clear()
        print("Feature cache cleared")


__all__ = ["FastRecommender", "ItemFeatureStore", "build_history_csr", "merge_topk"]

"""PxrEngine: torch-tensor wrapper over the libpxr C ABI (include/pxr.h).

PyTorch is plumbing here: it owns device memory and streams; every kernel that
touches a score is in libpxr.so.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import PxrError


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _dev_idx(t, device, dtype=torch.int64):
    """Index tensors reach the kernels as raw pointers: normalise dtype / device / layout here (no-ops when the caller
    already passes the right thing)."""
    return torch.as_tensor(t).detach().to(device=device, dtype=dtype).contiguous()


def check_index_range(idx, n: int, what: str):
    """The reference raises IndexError from ``nn.Embedding`` on an out-of-range index (multimodal.py:553-555); the
    kernels would read out of bounds instead.  One min/max reduction (a device sync for CUDA tensors): used on the
    API entry points that take caller-supplied indices, not inside the block loop."""
    t = torch.as_tensor(idx)
    if t.numel() == 0:
        return
    lo, hi = int(t.min()), int(t.max())
    if lo < 0 or hi >= n:
        raise IndexError(f"{what} out of range: [{lo}, {hi}] not inside [0, {n})")


class PxrEngine:
    """One handle = one model configuration on one GPU (not thread-safe)."""

    MAX_FUSED_K = 1024      # PXR_TC_MAX_K: top_k the fused path serves (64 list slots per pass, one pass per page of 64)

    def __init__(self, *, fusion_type: str, embedding_dim: int, vision_dim: int, language_dim: int,
                 num_numerical: int, hidden_dims: Sequence[int], n_tags: int, num_heads: int = 4,
                 activation: str = "relu", final_activation: str = "sigmoid", use_batch_norm: bool = True,
                 projection_hidden_dim: Optional[int] = None, path: str = "auto", precision: str = "bf16",
                 device: Optional[torch.device] = None, rescore: bool = True):
        if not torch.cuda.is_available():
            raise PxrError("PxrEngine needs a CUDA device: the scoring path has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise PxrError(f"PxrEngine needs a CUDA device, got {self.device}")
        if fusion_type not in _lib.FUSION:
            raise ValueError(f"Unknown fusion type: '{fusion_type}'")
        cfg = _lib.PxrConfig()
        cfg.struct_size = C.sizeof(_lib.PxrConfig)
        cfg.fusion = _lib.FUSION[fusion_type]
        cfg.embedding_dim = embedding_dim
        cfg.vision_dim = int(vision_dim or 0)
        cfg.language_dim = int(language_dim or 0)
        cfg.num_numerical = int(num_numerical or 0)
        cfg.projection_hidden = int(projection_hidden_dim or 0)
        if len(hidden_dims) > _lib.PXR_MAX_HIDDEN:
            raise ValueError(f"at most {_lib.PXR_MAX_HIDDEN} hidden layers are supported")
        cfg.n_hidden = len(hidden_dims)
        for i, hdim in enumerate(hidden_dims):
            cfg.hidden[i] = int(hdim)
        cfg.num_heads = num_heads
        # unknown activation names fall back to ReLU like the reference (multimodal.py:167)
        cfg.activation = _lib.ACT.get((activation or "relu").lower(), 0)
        cfg.final_activation = _lib.FINAL.get((final_activation or "none").lower(), 0)
        cfg.use_batch_norm = int(bool(use_batch_norm))
        cfg.n_tags = n_tags
        cfg.path = _lib.PATH[path]
        cfg.precision = _lib.PRECISION[precision]
        self.cfg = cfg
        self.fusion_type = fusion_type
        self.D = embedding_dim
        self.M = 3 + (cfg.vision_dim > 0) + (cfg.language_dim > 0) + (cfg.num_numerical > 0)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.pxr_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise PxrError(f"pxr_create failed ({rc}): {self.lib.pxr_last_error(None).decode()}")
        self.set_rescore(rescore)
        if path == "auto" and self.active_path != "tcgen05":
            # never silent: the generic kernels are a correctness path, two orders of magnitude slower
            logging.getLogger(__name__).warning(
                "PxrEngine: the fused tcgen05 kernel does not cover this model configuration (%s); scoring runs on the generic "
                "fp32 SIMT kernels (exact, ~100x slower)", self.path_reason)
        self._warned_k = False
        self._keep: Dict[str, object] = {}
        self._items_ws: Optional[torch.Tensor] = None
        self._score_ws: Optional[torch.Tensor] = None
        self.n_rows = 0
        self.item_base = 0
        self.items_token = None

    # ------------------------------------------------------------------ util
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.pxr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise PxrError(f"{what} failed ({rc}): {self.lib.pxr_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self.lib.pxr_launch_count(self._h))

    @property
    def active_path(self) -> str:
        return {1: "simt", 2: "tcgen05"}.get(self.lib.pxr_active_path(self._h), "?")

    @property
    def path_reason(self) -> str:
        """Why the generic SIMT kernels are active ('' on the fused path)."""
        return self.lib.pxr_path_reason(self._h).decode()

    def profile(self, on: bool = True):
        """Time the dominant kernel with CUDA events on the launching stream."""
        self._check(self.lib.pxr_profile_enable(self._h, int(on)), "pxr_profile_enable")

    def profile_read(self) -> Tuple[float, int]:
        """(summed ms, launches) of the dominant kernel since the last read."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        self._check(self.lib.pxr_profile_read(self._h, C.byref(ms), C.byref(n)), "pxr_profile_read")
        return float(ms.value), int(n.value)

    def set_rescore(self, on: bool):
        """Exact mode of the fused path (include/pxr.h: pxr_set_rescore): fp32 re-score + re-rank of the 64
        candidates the 16-bit kernel keeps per user.  On by default."""
        self._check(self.lib.pxr_set_rescore(self._h, int(bool(on))), "pxr_set_rescore")

    @property
    def rescore(self) -> bool:
        return bool(self.lib.pxr_get_rescore(self._h))

    def set_small_batch(self, mode: int):
        """Small-batch tile shape of the fused gated kernel (include/pxr.h: pxr_set_small_batch): -1 = when the cost
        model expects a gain (default), 0 = never, 1 = whenever the call allows it (<= 8 users, K <= 64)."""
        self._check(self.lib.pxr_set_small_batch(self._h, int(mode)), "pxr_set_small_batch")

    def set_records_only(self, on: bool = True):
        """Keep only the fp32 item records at the next ``precompute_items`` (a handle that serves ``rescore_topk`` /
        ``score_pairs`` only, e.g. the whole-catalogue re-score records of an item-sharded rank)."""
        self._check(self.lib.pxr_set_records_only(self._h, int(bool(on))), "pxr_set_records_only")

    def set_path(self, path: str):
        self._check(self.lib.pxr_set_path(self._h, _lib.PATH[path]), "pxr_set_path")

    # --------------------------------------------------------------- weights
    def load_weights(self, sd: Dict[str, torch.Tensor], use_batch_norm: bool, bn_eps: float = 1e-5):
        """``sd`` uses the reference state_dict key names."""
        dev = self.device
        keep: Dict[str, torch.Tensor] = {}

        def g(key):
            if key not in sd:
                return None
            keep[key] = _dev_f32(sd[key], dev)
            return keep[key]

        w = _lib.PxrWeights()
        w.struct_size = C.sizeof(_lib.PxrWeights)
        w.bn_eps = bn_eps

        def setp(field, key):
            t = g(key)
            setattr(w, field, t.data_ptr() if t is not None else None)

        setp("tag_embedding", "tag_embedding.weight")
        for name, pref in (("vision", "vision_projection"), ("language", "language_projection"),
                           ("numerical", "numerical_projection")):
            setp(f"{name}_w0", f"{pref}.0.weight")
            setp(f"{name}_b0", f"{pref}.0.bias")
            setp(f"{name}_w1", f"{pref}.3.weight")
            setp(f"{name}_b1", f"{pref}.3.bias")
        setp("gate_w", "fusion_layer.gating_network.0.weight")
        setp("gate_b", "fusion_layer.gating_network.0.bias")
        setp("attn_in_w", "fusion_layer.attention.in_proj_weight")
        setp("attn_in_b", "fusion_layer.attention.in_proj_bias")
        setp("attn_out_w", "fusion_layer.attention.out_proj.weight")
        setp("attn_out_b", "fusion_layer.attention.out_proj.bias")
        setp("attn_ln_w", "fusion_layer.norm.weight")
        setp("attn_ln_b", "fusion_layer.norm.bias")
        stride = 4 if use_batch_norm else 3
        n_hidden = self.cfg.n_hidden
        for l in range(n_hidden):
            base = l * stride
            for field, key in (("mlp_w", f"prediction_network.{base}.weight"),
                               ("mlp_b", f"prediction_network.{base}.bias")):
                t = g(key)
                getattr(w, field)[l] = t.data_ptr() if t is not None else None
            if use_batch_norm:
                for field, suf in (("bn_w", "weight"), ("bn_b", "bias"), ("bn_mean", "running_mean"),
                                   ("bn_var", "running_var")):
                    t = g(f"prediction_network.{base + 2}.{suf}")
                    getattr(w, field)[l] = t.data_ptr() if t is not None else None
        last = n_hidden * stride
        setp("out_w", f"prediction_network.{last}.weight")
        setp("out_b", f"prediction_network.{last}.bias")
        with torch.cuda.device(dev):
            self._check(self.lib.pxr_load_weights(self._h, C.byref(w), _stream()), "pxr_load_weights")
            torch.cuda.current_stream().synchronize()   # the fp32 sources in `keep` may now be dropped
        self.n_rows = 0
        self._items_ws = None

    # ----------------------------------------------------------------- items
    def precompute_items(self, item_embedding: torch.Tensor, tag_idx: torch.Tensor,
                         vis: Optional[torch.Tensor], txt: Optional[torch.Tensor], num: Optional[torch.Tensor],
                         item_idx: Optional[torch.Tensor] = None, item_base: int = 0,
                         n_rows: Optional[int] = None, validate: bool = True):
        """K1+K2 over ``n_rows`` item rows.  Row r describes global item
        ``item_idx[r]`` (or ``item_base + r``); feature tensors are row-aligned.
        ``validate=False`` skips the index range checks (two reductions and a device sync) for callers that vouch for
        their indices -- e.g. to capture the call in a CUDA graph."""
        dev = self.device
        n = int(n_rows if n_rows is not None else tag_idx.shape[0])
        emb = _dev_f32(item_embedding, dev)
        if emb.dim() != 2 or emb.shape[1] != self.D:
            raise ValueError(f"item_embedding must be (n_items, {self.D}), got {tuple(emb.shape)}")
        tag = _dev_idx(tag_idx, dev)
        if tag.shape[0] < n:
            raise ValueError(f"tag_idx has {tag.shape[0]} entries for {n} item rows")
        if validate:
            check_index_range(tag[:n], int(self.cfg.n_tags), "tag_idx")
        v = _dev_f32(vis, dev) if vis is not None and self.cfg.vision_dim else None
        t = _dev_f32(txt, dev) if txt is not None and self.cfg.language_dim else None
        x = _dev_f32(num, dev) if num is not None and self.cfg.num_numerical else None
        for name, ten, dim in (("vision", v, self.cfg.vision_dim), ("language", t, self.cfg.language_dim),
                               ("numerical", x, self.cfg.num_numerical)):
            if dim and (ten is None or ten.shape[0] < n or ten.shape[-1] != dim):
                raise ValueError(f"{name} features must be ({n}, {dim}), got "
                                 f"{None if ten is None else tuple(ten.shape)}")
        ii = _dev_idx(item_idx, dev) if item_idx is not None else None
        if ii is not None:
            if validate:
                check_index_range(ii[:n], int(emb.shape[0]), "item_idx")
        elif int(item_base) < 0 or int(item_base) + n > int(emb.shape[0]):
            raise IndexError(f"item rows [{item_base}, {int(item_base) + n}) not inside the item embedding table ({emb.shape[0]} rows)")
        nbytes = int(self.lib.pxr_items_bytes(self._h, n))
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 256
        with torch.cuda.device(dev):
            self._check(self.lib.pxr_precompute_items(
                self._h, _ptr(emb), _ptr(ii), _ptr(tag), _ptr(v), _ptr(t), _ptr(x), n, int(item_base),
                C.c_void_p(ws.data_ptr() + off), nbytes, _stream()), "pxr_precompute_items")
        self._items_ws = ws
        self._keep["items_in"] = (emb, tag, v, t, x, ii)   # inputs must outlive the async launch
        self._keep.pop("missing", None)
        self.n_rows, self.item_base = n, int(item_base)
        self.items_token = None                            # set by the owner of the records (FastRecommender.engine)

    def set_missing_items(self, flags: Optional[torch.Tensor]):
        """Per-row flags (bool / uint8, aligned with the precomputed rows): flagged items score exactly 0.0, as
        items without features do in the reference (recommender.py:229-230).  ``None`` clears."""
        if flags is None:
            self._check(self.lib.pxr_set_missing_items(self._h, None, 0), "pxr_set_missing_items")
            self._keep.pop("missing", None)
            return
        f = flags.to(device=self.device, dtype=torch.uint8).contiguous()
        self._check(self.lib.pxr_set_missing_items(self._h, _ptr(f), int(f.shape[0])), "pxr_set_missing_items")
        self._keep["missing"] = f

    # --------------------------------------------------------------- scoring
    def _user_args(self, user_embedding: torch.Tensor, user_idx):
        """dtype / device / layout of the two tensors every scoring call hands to the kernels as raw pointers
        (index RANGES are checked by the callers that receive them from outside: FastRecommender, forward)."""
        dev = self.device
        if not isinstance(user_embedding, torch.Tensor) or user_embedding.dim() != 2 or user_embedding.shape[1] != self.D:
            raise ValueError(f"user_embedding must be a (n_users, {self.D}) tensor")
        if user_embedding.device != dev or user_embedding.dtype != torch.float32 or not user_embedding.is_contiguous():
            user_embedding = _dev_f32(user_embedding, dev)
        user_idx = _dev_idx(user_idx, dev)
        if user_idx.dim() != 1:
            raise ValueError("user_idx must be 1-D")
        return user_embedding, user_idx

    def score_topk(self, user_embedding: torch.Tensor, user_idx: torch.Tensor, k: int,
                   seen_indptr: Optional[torch.Tensor] = None, seen_idx: Optional[torch.Tensor] = None
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
        dev = self.device
        user_embedding, user_idx = self._user_args(user_embedding, user_idx)
        if (seen_indptr is None) != (seen_idx is None):
            raise ValueError("seen_indptr and seen_idx go together")
        if seen_indptr is not None:
            seen_indptr, seen_idx = _dev_idx(seen_indptr, dev), _dev_idx(seen_idx, dev, torch.int32)
            if seen_indptr.shape[0] != user_idx.shape[0] + 1:
                raise ValueError("seen_indptr must have n_users + 1 entries")
        n = int(user_idx.shape[0])
        if k > self.MAX_FUSED_K and self.active_path == "tcgen05" and not self._warned_k:
            self._warned_k = True
            logging.getLogger(__name__).warning(
                "PxrEngine.score_topk: top_k=%d exceeds the %d candidates (16 pages of 64 list slots) the fused tcgen05 kernel "
                "keeps per user; this call runs on the generic fp32 SIMT kernels (exact, ~100x slower)", k, self.MAX_FUSED_K)
        out_s = torch.empty((n, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((n, k), dtype=torch.int32, device=dev)
        nbytes = int(self.lib.pxr_score_topk_bytes(self._h, n, k))
        if self._score_ws is None or self._score_ws.numel() < nbytes + 256:
            self._score_ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        ws = self._score_ws
        off = (-ws.data_ptr()) % 256
        with torch.cuda.device(dev):
            self._check(self.lib.pxr_score_topk(
                self._h, _ptr(user_embedding), _ptr(user_idx), n, _ptr(seen_indptr), _ptr(seen_idx), k,
                _ptr(out_s), _ptr(out_i), C.c_void_p(ws.data_ptr() + off), ws.numel() - off, _stream()),
                "pxr_score_topk")
        return out_s, out_i

    def rescore_topk(self, user_embedding: torch.Tensor, user_idx: torch.Tensor, cand_idx: torch.Tensor, k: int
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
        """The re-score step of exact mode on its own (``pxr_rescore_lists``): (n, L) candidate lists of GLOBAL item
        indices (-1 padded; L = 64 pages of the fused kernel's lists, a multiple of 64 up to 1 024) -> the K <= L best by the
        fp32 arithmetic of ``score_pairs`` against this engine's records."""
        dev = self.device
        user_embedding, user_idx = self._user_args(user_embedding, user_idx)
        cand = _dev_idx(cand_idx, dev, torch.int32)
        n = int(user_idx.shape[0])
        L = int(cand.shape[1]) if cand.dim() == 2 else -1
        if cand.dim() != 2 or cand.shape[0] != n or L < 64 or L % 64 or L > self.MAX_FUSED_K or k > L:
            raise ValueError(f"cand_idx must be ({n}, L) with L a multiple of 64 up to {self.MAX_FUSED_K} and top_k <= L, got {tuple(cand.shape)}, top_k={k}")
        out_s = torch.empty((n, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((n, k), dtype=torch.int32, device=dev)
        nbytes = int(self.lib.pxr_rescore_lists_bytes(n, L))
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 256
        with torch.cuda.device(dev):
            self._check(self.lib.pxr_rescore_lists(self._h, _ptr(user_embedding), _ptr(user_idx), n, _ptr(cand), L, int(k), _ptr(out_s),
                                                   _ptr(out_i), C.c_void_p(ws.data_ptr() + off), ws.numel() - off, _stream()),
                        "pxr_rescore_lists")
        return out_s, out_i

    def score_pairs(self, user_embedding: torch.Tensor, user_idx: torch.Tensor, item_row: torch.Tensor,
                    want_logit: bool = False):
        dev = self.device
        user_embedding, user_idx = self._user_args(user_embedding, user_idx)
        item_row = _dev_idx(item_row, dev)
        if item_row.shape != user_idx.shape:
            raise ValueError("user_idx and item_row must have the same length")
        n = int(user_idx.shape[0])
        out = torch.empty(n, dtype=torch.float32, device=dev)
        logit = torch.empty(n, dtype=torch.float32, device=dev) if want_logit else None
        with torch.cuda.device(dev):
            self._check(self.lib.pxr_score_pairs(self._h, _ptr(user_embedding), _ptr(user_idx), _ptr(item_row), n,
                                                 _ptr(out), _ptr(logit), _stream()), "pxr_score_pairs")
        return (out, logit) if want_logit else out

    def topk_rows(self, scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Exact top-K of every row of a dense (n_rows, n_cols) fp32 score matrix (-inf = masked, ties -> lower
        column): (scores, column positions), padded with -inf / -1."""
        dev = self.device
        scores = scores.to(device=dev, dtype=torch.float32).contiguous()
        n, c = scores.shape
        out_s = torch.empty((n, k), dtype=torch.float32, device=dev)
        out_p = torch.empty((n, k), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            self._check(self.lib.pxr_topk_rows(self._h, _ptr(scores), n, c, k, _ptr(out_s), _ptr(out_p), _stream()),
                        "pxr_topk_rows")
        return out_s, out_p

    # ---------------------------------------------------------------- merges
    def merge_topk(self, scores: torch.Tensor, idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(S, n_users, K) per-shard lists -> (n_users, K); ties -> lower global index."""
        return merge_topk(scores, idx)


def merge_topk(scores: torch.Tensor, idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    S, n, k = scores.shape
    scores = scores.contiguous()
    idx = idx.to(torch.int32).contiguous()
    out_s = torch.empty((n, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((n, k), dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        rc = lib.pxr_merge_topk(_ptr(scores), _ptr(idx), S, n, k, _ptr(out_s), _ptr(out_i), _stream())
    if rc != 0:
        raise PxrError(f"pxr_merge_topk failed ({rc})")
    return out_s, out_i


_METRIC_COLS = ("precision", "recall", "f1", "hit_rate", "ndcg", "mrr", "ndcg_list_ideal", "precision_hits_over_k", "map")


def sample_candidates(user_idx: torch.Tensor, pos_indptr: torch.Tensor, pos_idx: torch.Tensor, n_items: int,
                      n_neg: int, seed: int, stride: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Candidate lists of the sampled protocol (pxr_sample_candidates): per user the positives
    (``pos_idx`` ascending inside each user) plus ``n_neg`` uniformly sampled negatives, shuffled;
    returns ((n_users, stride) int32 item indices, -1 padded, (n_users,) int32 lengths)."""
    lib = _lib.load()
    dev = pos_idx.device
    n = int(pos_indptr.shape[0]) - 1
    pos_indptr = pos_indptr.to(device=dev, dtype=torch.int64).contiguous()
    pos_idx = pos_idx.to(device=dev, dtype=torch.int32).contiguous()
    if stride is None:
        max_pos = int((pos_indptr[1:] - pos_indptr[:-1]).max().item()) if n else 0
        stride = max(1, min(1024, max_pos + int(n_neg)))
    uidx = user_idx.to(device=dev, dtype=torch.int64).contiguous() if user_idx is not None else None
    cand = torch.empty((n, stride), dtype=torch.int32, device=dev)
    length = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pxr_sample_candidates(_ptr(uidx), n, _ptr(pos_indptr), _ptr(pos_idx), int(n_items), int(n_neg),
                                       C.c_uint64(int(seed) & ((1 << 64) - 1)), int(stride), _ptr(cand), _ptr(length),
                                       _stream())
    if rc != 0:
        raise PxrError(f"pxr_sample_candidates failed ({rc})")
    return cand, length


def weighted_candidates(user_idx: torch.Tensor, pos_indptr: torch.Tensor, pos_idx: torch.Tensor, weights: torch.Tensor,
                        n_neg: int, seed: int, stride: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Candidate lists of the popularity-biased strategies (pxr_weighted_candidates; reference
    src/evaluation/tasks.py:225-308): per user the positives plus ``n_neg`` negatives drawn without replacement with
    probability proportional to ``weights`` (float64, one per item), shuffled like ``sample_candidates``.  A pure
    function of (seed, user index).  ``weights`` must live on the GPU: there is no host path."""
    lib = _lib.load()
    dev = weights.device
    if dev.type != "cuda":
        raise PxrError("weighted_candidates runs on the GPU only (pxr_weighted_candidates): pass CUDA tensors")
    n = int(pos_indptr.shape[0]) - 1
    n_items = int(weights.shape[0])
    w = weights.to(dtype=torch.float64).contiguous()
    pos_indptr = pos_indptr.to(device=dev, dtype=torch.int64).contiguous()
    pos_idx = pos_idx.to(device=dev, dtype=torch.int32).contiguous()
    if stride is None:
        max_pos = int((pos_indptr[1:] - pos_indptr[:-1]).max().item()) if n else 0
        stride = max(1, min(1024, max_pos + int(n_neg)))
    uidx = user_idx.to(device=dev, dtype=torch.int64).contiguous() if user_idx is not None else None
    cand = torch.empty((n, stride), dtype=torch.int32, device=dev)
    length = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pxr_weighted_candidates(_ptr(uidx), n, _ptr(pos_indptr), _ptr(pos_idx), _ptr(w), n_items, int(n_neg),
                                         C.c_uint64(int(seed) & ((1 << 64) - 1)), int(stride), _ptr(cand), _ptr(length),
                                         _stream())
    if rc != 0:
        raise PxrError(f"pxr_weighted_candidates failed ({rc})")
    return cand, length


_METRIC_TABLES = {}


def _metric_tables(kstride: int, dev):
    """discount[j] = 1/log2(j+2) and its running sums, computed with numpy / Python left-to-right adds so the device
    results are bit-identical to the reference's numpy arithmetic (tasks.py:733-747); cached per (K, device)."""
    key = (kstride, str(dev))
    if key not in _METRIC_TABLES:
        disc = np.array([1.0 / np.log2(i + 2) for i in range(kstride)], dtype=np.float64)
        ideal = np.zeros(kstride + 1, dtype=np.float64)
        acc = 0
        for i in range(kstride):
            acc = acc + disc[i]          # Python left-to-right sum, as in tasks.py:744
            ideal[i + 1] = acc
        _METRIC_TABLES[key] = (torch.from_numpy(disc).to(dev), torch.from_numpy(ideal).to(dev))
    return _METRIC_TABLES[key]


def ranking_metric_sums(topk_idx: torch.Tensor, gt_indptr: torch.Tensor, gt_idx: torch.Tensor,
                        ks: Sequence[int], as_device: bool = False, recall_den: Optional[torch.Tensor] = None):
    """K5: per-cut-off sums over users, shape (len(ks), 9) float64 (host; the device tensor, without a
    synchronisation, when ``as_device``).  Column order: precision (hits / len(recs)), recall, f1, hit_rate, ndcg
    (tasks.py), mrr, ndcg (metrics.py), precision (hits / k, metrics.py:35), average precision (metrics.py:102-133).
    ``recall_den``: per-user recall denominators (raw number of test rows, tasks.py:579); default = the number of
    distinct positives in the CSR."""
    lib = _lib.load()
    dev = topk_idx.device
    topk_idx = topk_idx.to(torch.int32).contiguous()
    n, kstride = topk_idx.shape
    ks = sorted(int(k) for k in ks)
    d_disc, d_ideal = _metric_tables(kstride, dev)
    out = torch.empty((len(ks), _lib.PXR_METRIC_COLS), dtype=torch.float64, device=dev)
    rden = recall_den.to(device=dev, dtype=torch.int32).contiguous() if recall_den is not None else None
    if rden is not None and rden.shape[0] != n:
        raise ValueError("recall_den must have one entry per user")
    nbytes = int(lib.pxr_metrics_bytes(n, len(ks)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ks_arr = (C.c_int32 * len(ks))(*ks)
    gt_indptr = gt_indptr.to(device=dev, dtype=torch.int64).contiguous()
    gt_idx = gt_idx.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        rc = lib.pxr_metrics(_ptr(topk_idx), kstride, n, _ptr(gt_indptr), _ptr(gt_idx), _ptr(rden), ks_arr, len(ks),
                             _ptr(d_disc), _ptr(d_ideal), _ptr(out), _ptr(ws), nbytes, _stream())
    if rc != 0:
        raise PxrError(f"pxr_metrics failed ({rc})")
    return out if as_device else out.cpu().numpy()

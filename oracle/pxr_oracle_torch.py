"""CPU arm of bench.py: the reference forward restated with the SAME library the
reference uses for its arithmetic (PyTorch eager, fp32, CPU, intra-op threads =
all host cores), so that the timed CPU baseline runs at the reference's own
speed rather than at numpy's.

TEST INFRASTRUCTURE ONLY (see oracle/pxr_oracle.py for the rules): imported by
``tests/`` (pinned there against tests/golden, i.e. against outputs of the
unmodified reference) and by ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs.  Never imported by the product package.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

_ACT = {"relu": F.relu, "gelu": F.gelu, "tanh": torch.tanh, "leaky_relu": lambda x: F.leaky_relu(x, 0.01),
        "silu": F.silu}


def _act(name):
    """src/models/multimodal.py:150-167 (unknown names -> ReLU)."""
    return _ACT.get((name or "relu").lower(), F.relu)


def _proj(x, sd, prefix, act):
    """src/models/multimodal.py:262-313."""
    y = act(F.linear(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"]))
    if prefix + ".3.weight" in sd:
        y = act(F.linear(y, sd[prefix + ".3.weight"], sd[prefix + ".3.bias"]))
    return y


@torch.no_grad()
def forward_pairs(sd: Dict[str, torch.Tensor], cfg: dict, user_idx, item_idx, tag_idx, vis=None, txt=None, num=None,
                  return_logit: bool = False) -> torch.Tensor:
    """MultimodalRecommender.forward on B pairs (src/models/multimodal.py:528-610)."""
    act = _act(cfg.get("fusion_activation", "relu"))
    feats = [F.embedding(user_idx, sd["user_embedding.weight"]), F.embedding(item_idx, sd["item_embedding.weight"]),
             F.embedding(tag_idx, sd["tag_embedding.weight"])]                     # :553-555
    if vis is not None and "vision_projection.0.weight" in sd:
        feats.append(_proj(vis, sd, "vision_projection", act))                     # :559-561
    if txt is not None and "language_projection.0.weight" in sd:
        feats.append(_proj(txt, sd, "language_projection", act))                   # :564-566
    if num is not None and "numerical_projection.0.weight" in sd:
        feats.append(_proj(num, sd, "numerical_projection", act))                  # :569-570
    ft = cfg.get("fusion_type", "concatenate")
    if ft == "concatenate":
        fused = torch.cat(feats, dim=1)                                            # :583-584
    elif ft == "gated":                                                            # src/models/layers.py:195-225
        g = torch.softmax(F.linear(torch.cat(feats, dim=1), sd["fusion_layer.gating_network.0.weight"],
                                   sd["fusion_layer.gating_network.0.bias"]), dim=-1)
        fused = (torch.stack(feats, dim=1) * g.unsqueeze(-1)).sum(dim=1)
    elif ft == "attention":                                                        # src/models/layers.py:135-164
        x = torch.stack(feats, dim=0)                                              # (M, B, D)
        D = x.shape[-1]
        a, _ = F.multi_head_attention_forward(
            x, x, x, D, int(cfg.get("num_attention_heads", 4)), sd["fusion_layer.attention.in_proj_weight"],
            sd["fusion_layer.attention.in_proj_bias"], None, None, False, 0.0,
            sd["fusion_layer.attention.out_proj.weight"], sd["fusion_layer.attention.out_proj.bias"],
            training=False, need_weights=False)
        z = F.layer_norm(x + a, (D,), sd["fusion_layer.norm.weight"], sd["fusion_layer.norm.bias"], 1e-5)
        fused = z.mean(dim=0)
    else:
        raise ValueError(f"Unknown fusion type: '{ft}'")
    # prediction network: [Linear -> act -> BatchNorm1d(eval) -> Dropout(id)] x L -> Linear (:366-386, 594)
    idxs = sorted({int(k.split(".")[1]) for k in sd if k.startswith("prediction_network.") and k.endswith(".weight")
                   and sd[k].dim() == 2})
    h = fused
    for i in idxs[:-1]:
        p = f"prediction_network.{i}"
        h = act(F.linear(h, sd[p + ".weight"], sd[p + ".bias"]))
        if cfg.get("use_batch_norm", True):
            bn = f"prediction_network.{i + 2}"
            h = F.batch_norm(h, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                             False, 0.0, 1e-5)
    p = f"prediction_network.{idxs[-1]}"
    z = F.linear(h, sd[p + ".weight"], sd[p + ".bias"])[:, 0]
    if return_logit:
        return z
    fin = (cfg.get("final_activation") or "none").lower()
    y = torch.sigmoid(z) if fin == "sigmoid" else torch.tanh(z) if fin == "tanh" else z
    return torch.nan_to_num(y, nan=0.0, posinf=10.0, neginf=-10.0)                 # :596-597


@torch.no_grad()
def recommend_block(sd, cfg, users: torch.Tensor, feats: Dict[str, torch.Tensor], k: int,
                    seen_indptr: Optional[torch.Tensor] = None, seen_idx: Optional[torch.Tensor] = None,
                    chunk_pairs: int = 1 << 18):
    """Every user of ``users`` against every item: the reference forward on
    flattened (user, item) pairs in chunks, seen items dropped, stable
    descending order, first k (src/inference/recommender.py:73-106).  Returns
    (idx (n, k) int64 padded with -1, scores (n, k) padded with -inf)."""
    NI = int(feats["tag_idx"].shape[0])
    items = torch.arange(NI)
    out_i = torch.full((len(users), k), -1, dtype=torch.int64)
    out_s = torch.full((len(users), k), float("-inf"))
    per = max(1, chunk_pairs // NI)
    for u0 in range(0, len(users), per):
        uu = users[u0:u0 + per]
        ui = uu.repeat_interleave(NI)
        ii = items.repeat(len(uu))
        s = forward_pairs(sd, cfg, ui, ii, feats["tag_idx"][ii], feats["vis"][ii] if "vis" in feats else None,
                          feats["txt"][ii] if "txt" in feats else None, feats["num"][ii] if "num" in feats else None)
        s = s.view(len(uu), NI).clone()
        for r, u in enumerate(uu.tolist()):
            if seen_indptr is not None:
                s[r, seen_idx[int(seen_indptr[u]):int(seen_indptr[u + 1])].long()] = float("-inf")
            order = torch.sort(s[r], descending=True, stable=True).indices[:k]
            keep = order[s[r][order] > float("-inf")]            # real scores are finite after the NaN/Inf guard
            out_i[u0 + r, :len(keep)] = keep
            out_s[u0 + r, :len(keep)] = s[r][keep]
    return out_i, out_s


# --------------------------------------------------------------------------
# The LITERAL per-user path of the reference (what a reference user experiences): one get_recommendations call per
# user, 256-item batches, per-item feature-dict fetches, torch.stack per key, sklearn LabelEncoder.transform per batch,
# one forward per batch, .cpu().tolist(), Python sort.  Restated step by step so bench.py can time it beside the
# batched forward; pinned in tests/ to the lists the real reference Recommender produced
# (tests/golden/recommender_lists.json).
# --------------------------------------------------------------------------
class LiteralRecommender:
    """src/inference/recommender.py:30-110, 144-269 with the model call replaced by ``forward_pairs`` above (the same
    arithmetic, pinned to the reference module's outputs).  ``dataset`` is duck-typed like the reference's:
    ``user_encoder`` / ``item_encoder`` (sklearn LabelEncoder), ``feature_cache`` (dict-like ``get``),
    ``get_user_history(user_id) -> set``."""

    def __init__(self, sd, cfg, dataset):
        self.sd, self.cfg, self.dataset = sd, cfg, dataset

    def _get_item_features(self, item_id_str):                       # recommender.py:239-246 (cache hit path)
        item_id_str = str(item_id_str)
        if self.dataset.feature_cache and self.dataset.feature_cache.get(item_id_str):
            return self.dataset.feature_cache.get(item_id_str)
        return None

    @torch.no_grad()
    def _score_items_batch(self, user_tensor, item_ids_str):         # recommender.py:144-236
        if len(item_ids_str) == 0:
            return []
        item_features_list, valid_ids = [], []
        for item_id in item_ids_str:                                  # :162-166
            features = self._get_item_features(item_id)
            if features:
                item_features_list.append(features)
                valid_ids.append(item_id)
        if not valid_ids:
            return [0.0] * len(item_ids_str)
        keys = item_features_list[0].keys()
        collated = {key: torch.stack([d[key] for d in item_features_list]) for key in keys}          # :175-178
        user_batch = user_tensor.repeat(len(valid_ids))                                              # :182
        item_idx = torch.tensor(self.dataset.item_encoder.transform(valid_ids), dtype=torch.long)   # :191
        scores = forward_pairs(self.sd, self.cfg, user_batch, item_idx, collated["tag_idx"], collated.get("image"),
                               collated.get("text_input_ids"), collated.get("numerical_features"))   # :221-222
        scores = scores.squeeze().cpu().tolist()                                                     # :224
        if not isinstance(scores, list):
            scores = [scores]
        final = {i: s for i, s in zip(valid_ids, scores)}                                            # :229-230
        return [final.get(i, 0.0) for i in item_ids_str]

    def get_recommendations(self, user_id, top_k=10, filter_seen=True, candidates=None):            # recommender.py:52-110
        user_id = str(user_id)
        user_classes = [str(c) for c in self.dataset.user_encoder.classes_]                          # :64 (O(n_users) per call)
        if user_id not in user_classes:
            return []
        user_encoded = self.dataset.user_encoder.transform([user_id])[0]                             # :69
        user_tensor = torch.tensor([user_encoded], dtype=torch.long)
        if candidates is None:
            cand = [str(i) for i in self.dataset.item_encoder.classes_]                              # :76
        else:
            item_classes = [str(c) for c in self.dataset.item_encoder.classes_]
            cand = [str(i) for i in candidates if str(i) in item_classes]                            # :81-82
        if not cand:
            return []
        if filter_seen:
            seen = self.dataset.get_user_history(user_id)                                            # :88-90
            cand = [i for i in cand if i not in seen]
        if not cand:
            return []
        item_scores = []
        for i in range(0, len(cand), 256):                                                           # :97-103
            batch = cand[i:i + 256]
            for item_id, sc in zip(batch, self._score_items_batch(user_tensor, batch)):
                item_scores.append((item_id, sc))
        item_scores.sort(key=lambda x: x[1], reverse=True)                                           # :105 (stable)
        return item_scores[:top_k]

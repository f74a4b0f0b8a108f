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

"""CPU oracle for the full-catalogue scoring / ranking / metrics path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``pixelrec_multimodal_b200/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker / the CPU arm that is timed beside the GPU path.

This is a plain-numpy restatement of the reference algorithm.  The arithmetic
of the reference lives in third-party PyTorch (``torch>=2.2.1`` unpinned in the
reference's ``setup.py:8``; CI pins 2.6.0; this container has 2.11.0), so the
restatement is pinned by running the *reference module itself* in this
container (stub identity backbones) and committing its outputs
as ``tests/golden/*.npz`` (generator: ``oracle/make_golden.py``).  The
reference's own known-answer tests for the metric formulas and the recommender
ordering semantics are restated in ``tests/test_oracle_golden.py``.

Every function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

try:  # erf for exact GELU (nn.GELU default, src/models/multimodal.py:160)
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover - scipy is in the image
    _erf = np.vectorize(math.erf)


# --------------------------------------------------------------------------
# elementary layers
# --------------------------------------------------------------------------
def activation(x: np.ndarray, name: str) -> np.ndarray:
    """src/models/multimodal.py:150-167 (unknown names fall back to ReLU)."""
    name = (name or "relu").lower()
    if name == "gelu":
        return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))
    if name == "tanh":
        return np.tanh(x)
    if name == "leaky_relu":
        return np.where(x >= 0, x, 0.01 * x)
    if name == "silu":
        return x / (1.0 + np.exp(-x))
    return np.maximum(x, 0.0)


def linear(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray]) -> np.ndarray:
    y = x @ w.T
    return y if b is None else y + b


def _projection(x, sd, prefix, act, dt):
    """src/models/multimodal.py:262-313: Linear->act->Dropout(id) or the
    two-layer variant when ``projection_hidden_dim`` is set (keys .0 and .3)."""
    y = activation(linear(x, sd[prefix + ".0.weight"].astype(dt), sd[prefix + ".0.bias"].astype(dt)), act)
    if prefix + ".3.weight" in sd:
        y = activation(linear(y, sd[prefix + ".3.weight"].astype(dt), sd[prefix + ".3.bias"].astype(dt)), act)
    return y


def modality_features(sd, act, user_idx, item_idx, tag_idx, vis=None, txt=None, num=None,
                      dtype=np.float64) -> List[np.ndarray]:
    """Feature list in the fixed order of src/models/multimodal.py:553-570."""
    dt = dtype
    feats = [sd["user_embedding.weight"].astype(dt)[user_idx],
             sd["item_embedding.weight"].astype(dt)[item_idx],
             sd["tag_embedding.weight"].astype(dt)[tag_idx]]
    if vis is not None and "vision_projection.0.weight" in sd:
        feats.append(_projection(vis.astype(dt), sd, "vision_projection", act, dt))
    if txt is not None and "language_projection.0.weight" in sd:
        feats.append(_projection(txt.astype(dt), sd, "language_projection", act, dt))
    if num is not None and "numerical_projection.0.weight" in sd:
        feats.append(_projection(num.astype(dt), sd, "numerical_projection", act, dt))
    return feats


def gated_fusion(sd, feats, dt, return_gates=False):
    """src/models/layers.py:195-225."""
    cat = np.concatenate(feats, axis=1)
    logits = linear(cat, sd["fusion_layer.gating_network.0.weight"].astype(dt),
                    sd["fusion_layer.gating_network.0.bias"].astype(dt))
    logits = logits - logits.max(axis=-1, keepdims=True)
    g = np.exp(logits)
    g /= g.sum(axis=-1, keepdims=True)
    stack = np.stack(feats, axis=1)                    # (B, M, D)
    if return_gates:
        return g
    return (stack * g[:, :, None]).sum(axis=1)


def attention_fusion(sd, feats, num_heads, dt, return_token_sum=False):
    """src/models/layers.py:135-164 as documented (list -> stack dim 0 ->
    nn.MultiheadAttention(batch_first=False) -> residual + LayerNorm(eps 1e-5)
    -> mean over tokens).  The call at src/models/multimodal.py:513-519 hands a
    tensor instead of a list and raises; SURVEY.md fact 3."""
    x = np.stack(feats, axis=0)                         # (M, B, D)
    M, B, D = x.shape
    w_in = sd["fusion_layer.attention.in_proj_weight"].astype(dt)
    b_in = sd["fusion_layer.attention.in_proj_bias"].astype(dt)
    qkv = x @ w_in.T + b_in                             # (M, B, 3D)
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    dh = D // num_heads
    q = q.reshape(M, B, num_heads, dh)
    k = k.reshape(M, B, num_heads, dh)
    v = v.reshape(M, B, num_heads, dh)
    s = np.einsum("abhd,cbhd->bhac", q, k) / math.sqrt(dh)   # (B, h, M, M)
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(axis=-1, keepdims=True)
    o = np.einsum("bhac,cbhd->abhd", p, v).reshape(M, B, D)
    o = o @ sd["fusion_layer.attention.out_proj.weight"].astype(dt).T + \
        sd["fusion_layer.attention.out_proj.bias"].astype(dt)
    y = x + o
    mu = y.mean(axis=-1, keepdims=True)
    var = ((y - mu) ** 2).mean(axis=-1, keepdims=True)
    z = (y - mu) / np.sqrt(var + 1e-5)
    if return_token_sum:        # sum over tokens of the normalised rows, before the LayerNorm affine and the 1/M
        return z.sum(axis=0)
    z = z * sd["fusion_layer.norm.weight"].astype(dt) + sd["fusion_layer.norm.bias"].astype(dt)
    return z.mean(axis=0)


def attention_token_sum_mma(sd, feats, num_heads, rnd=None, dt=np.float64):
    """The attention fusion of src/models/layers.py:135-164 in the form the fused kernel's register-MMA front end
    evaluates it (csrc/score_tc.cu: attn_item_step), with that kernel's operand roundings ``rnd`` (None = exact,
    then equal to attention_fusion(return_token_sum=True)).

    With P = I - 11^T/D (centring over d), Wc = P W_o, xc_a = P (x_a + b_o) and o_a the concatenated per-head
    attention outputs of row a, LayerNorm only needs yc_a = xc_a + Wc o_a.  Item rows a >= 1: the user column gets
    the weight w_ah = sigmoid(s_a0,h - L_ah) (L_ah = logsumexp of the item-item scores, per item), so
        yc_a = xc_a + sum_h (1 - w_ah) Nc_ah + Wc (w_a (.) v_u),     Nc_ah = Wc[:, head h] vbar_ah  (per item)
    and the user row:  yc_0 = xc_0 + Wc (p_00 (.) v_u) + sum_{h,b} p_0b,h Uc_bh,   Uc_bh = Wc[:, head h] v_b,h.
    Every product "coefficients x vectors" is a 16-bit MMA with fp32 accumulation: operands q / k (scores), w v_u,
    p_00 v_u, 1 - w, p_0b, Wc, Uc are rounded once; Nc and xc are carried as hi + lo pairs (two K rows of the same
    MMA), so they are exact to 2^-17.  Returns sum_a yc_a rstd_a (before the LayerNorm affine and the 1/M)."""
    r = rnd or (lambda a: a)
    M = len(feats)
    B, D = feats[0].shape
    dh = D // num_heads
    H = num_heads
    w_in = sd["fusion_layer.attention.in_proj_weight"].astype(dt)
    b_in = sd["fusion_layer.attention.in_proj_bias"].astype(dt)
    w_o = sd["fusion_layer.attention.out_proj.weight"].astype(dt)
    b_o = sd["fusion_layer.attention.out_proj.bias"].astype(dt)
    scale = 1.0 / math.sqrt(dh)
    centre = lambda a: a - a.mean(axis=-1, keepdims=True)
    Wc = w_o - w_o.mean(axis=0, keepdims=True)                   # P W_o: every column centred over the output index d
    Wc16 = r(Wc)
    hi_lo = (lambda a: (r(a), r(a - r(a)))) if rnd else (lambda a: (a, np.zeros_like(a)))

    def qkv(x):
        y = x @ w_in.T + b_in
        return y[:, :D] * scale, y[:, D:2 * D], y[:, 2 * D:]

    heads = lambda a: a.reshape(B, H, dh)
    eu = feats[0]
    qu, ku, vu = qkv(eu)
    s00 = (heads(qu) * heads(ku)).sum(2)                          # fp32 on CUDA cores, unrounded operands
    xc0 = centre(eu + b_o)
    qu16, ku16, vu16 = r(qu), r(ku), r(vu)
    nt = M - 1
    xs = feats[1:]
    q, k, v = zip(*[qkv(x) for x in xs])
    # ---- per-item record (item_attn_kernel): exact fp32 item-item softmax, then the stored 16-bit operands
    rec = []
    for a in range(nt):
        s = np.stack([(heads(q[a]) * heads(k[b])).sum(2) for b in range(nt)], 2)      # (B, H, nt)
        m = s.max(2, keepdims=True)
        e = np.exp(s - m)
        L = (m + np.log(e.sum(2, keepdims=True)))[:, :, 0]
        pi = e / e.sum(2, keepdims=True)
        vbar = sum(pi[:, :, b:b + 1] * heads(v[b]) for b in range(nt))                 # (B, H, dh)
        Nc = np.stack([vbar[:, h] @ Wc[:, h * dh:(h + 1) * dh].T for h in range(H)], 1)   # (B, H, D)
        Uc = np.stack([heads(v[a])[:, h] @ Wc[:, h * dh:(h + 1) * dh].T for h in range(H)], 1)
        rec.append(dict(L=L, q16=r(q[a]), k16=r(k[a]), Nc=hi_lo(Nc), xc=hi_lo(centre(xs[a] + b_o)), Uc16=r(Uc)))
    rstd = lambda y: 1.0 / np.sqrt((y * y).mean(1, keepdims=True) + 1e-5)
    per_head = lambda c: np.repeat(c, dh, axis=1)                  # (B, H) -> (B, D)
    # ---- user row
    S0 = np.stack([(heads(qu16) * heads(rec[b]["k16"])).sum(2) for b in range(nt)], 2)   # (B, H, nt)
    m = np.maximum(s00, S0.max(2))
    e0 = np.exp(s00 - m); eb = np.exp(S0 - m[:, :, None])
    inv = 1.0 / (e0 + eb.sum(2))
    p00, p0b = r(e0 * inv), r(eb * inv[:, :, None])
    y0 = xc0 + r(per_head(p00) * vu16) @ Wc16.T
    for b in range(nt):
        y0 = y0 + (p0b[:, :, b][:, :, None] * rec[b]["Uc16"]).sum(1)
    acc = y0 * rstd(y0)
    # ---- item rows
    for a in range(nt):
        sa0 = (heads(rec[a]["q16"]) * heads(ku16)).sum(2)
        w = r(1.0 / (1.0 + np.exp(rec[a]["L"] - sa0)))          # the weight is rounded once, both uses start from it
        omw = r(1.0 - w)
        y = rec[a]["xc"][0] + rec[a]["xc"][1] + (omw[:, :, None] * (rec[a]["Nc"][0] + rec[a]["Nc"][1])).sum(1) + \
            r(per_head(w) * vu16) @ Wc16.T
        acc = acc + y * rstd(y)
    return acc


def prediction_layout(sd, use_batch_norm: bool) -> Tuple[List[int], int]:
    """Indices of the hidden Linears and of the output Linear inside
    ``prediction_network`` (src/models/multimodal.py:371-386)."""
    idxs = sorted({int(k.split(".")[1]) for k in sd if k.startswith("prediction_network.") and
                   k.endswith(".weight") and sd[k].ndim == 2})
    return idxs[:-1], idxs[-1]


def prediction_network(sd, x, act, use_batch_norm, dt, return_logit=False):
    """[Linear -> act -> BatchNorm1d(eval, eps 1e-5) -> Dropout(id)] x L ->
    Linear(H_L, 1) (src/models/multimodal.py:366-386, 594)."""
    hidden, last = prediction_layout(sd, use_batch_norm)
    for i in hidden:
        p = f"prediction_network.{i}"
        x = activation(linear(x, sd[p + ".weight"].astype(dt), sd[p + ".bias"].astype(dt)), act)
        if use_batch_norm:
            bn = f"prediction_network.{i + 2}"
            x = (x - sd[bn + ".running_mean"].astype(dt)) / np.sqrt(sd[bn + ".running_var"].astype(dt) + 1e-5)
            x = x * sd[bn + ".weight"].astype(dt) + sd[bn + ".bias"].astype(dt)
    p = f"prediction_network.{last}"
    return linear(x, sd[p + ".weight"].astype(dt), sd[p + ".bias"].astype(dt))


def final_activation(z: np.ndarray, name: str) -> np.ndarray:
    """Sigmoid / Tanh / none (multimodal.py:381-384) then the NaN/Inf guard
    (multimodal.py:596-597: nan->0, +inf->10, -inf->-10)."""
    name = (name or "none").lower()
    with np.errstate(over="ignore"):
        if name == "sigmoid":
            y = 1.0 / (1.0 + np.exp(-z))
        elif name == "tanh":
            y = np.tanh(z)
        else:
            y = z
    return np.nan_to_num(y, nan=0.0, posinf=10.0, neginf=-10.0)


def forward_pairs(sd: Dict[str, np.ndarray], cfg: dict, user_idx, item_idx, tag_idx,
                  vis=None, txt=None, num=None, dtype=np.float64, return_logit=False) -> np.ndarray:
    """``MultimodalRecommender.forward`` on B pairs -> (B,) scores
    (src/models/multimodal.py:528-610).  ``cfg`` keys: fusion_type,
    fusion_activation, use_batch_norm, final_activation, num_attention_heads."""
    act = cfg.get("fusion_activation", "relu")
    feats = modality_features(sd, act, np.asarray(user_idx), np.asarray(item_idx), np.asarray(tag_idx),
                              vis, txt, num, dtype)
    ft = cfg.get("fusion_type", "concatenate")
    if ft == "concatenate":
        fused = np.concatenate(feats, axis=1)
    elif ft == "gated":
        fused = gated_fusion(sd, feats, dtype)
    elif ft == "attention":
        fused = attention_fusion(sd, feats, int(cfg.get("num_attention_heads", 4)), dtype)
    else:
        raise ValueError(f"Unknown fusion type: '{ft}'")
    z = prediction_network(sd, fused, act, bool(cfg.get("use_batch_norm", True)), dtype)[:, 0]
    if return_logit:
        return z
    return final_activation(z, cfg.get("final_activation", "sigmoid"))


def score_block(sd, cfg, user_indices: np.ndarray, item_lo: int, item_hi: int, feats: dict,
                dtype=np.float64, return_logit=False, chunk_pairs: int = 1 << 18) -> np.ndarray:
    """Dense score matrix [len(user_indices), item_hi-item_lo]: every user
    against every item, i.e. what the loop of
    src/inference/recommender.py:97-103 produces with ``candidates=None``."""
    users = np.asarray(user_indices, dtype=np.int64)
    items = np.arange(item_lo, item_hi, dtype=np.int64)
    out = np.empty((len(users), len(items)), dtype=dtype)
    per = max(1, chunk_pairs // max(1, len(items)))
    for u0 in range(0, len(users), per):
        uu = users[u0:u0 + per]
        ui = np.repeat(uu, len(items))
        ii = np.tile(items, len(uu))
        s = forward_pairs(sd, cfg, ui, ii, feats["tag_idx"][ii],
                          feats.get("vis")[ii] if feats.get("vis") is not None else None,
                          feats.get("txt")[ii] if feats.get("txt") is not None else None,
                          feats.get("num")[ii] if feats.get("num") is not None else None,
                          dtype=dtype, return_logit=return_logit)
        out[u0:u0 + len(uu)] = s.reshape(len(uu), len(items))
    return out


# --------------------------------------------------------------------------
# ranking (Recommender.get_recommendations semantics)
# --------------------------------------------------------------------------
def topk_from_scores(scores: np.ndarray, k: int, seen: Optional[Iterable[int]] = None,
                     item_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Candidates in index order, seen items dropped, *stable* descending sort,
    first k (src/inference/recommender.py:73-106).  Stable sort over
    index-ordered candidates means ties go to the lower item index.  Returns
    (global item indices int64, scores), length <= k."""
    n = scores.shape[0]
    cand = np.arange(n, dtype=np.int64)
    if seen is not None:
        seen = np.asarray(list(seen), dtype=np.int64) - item_base
        seen = seen[(seen >= 0) & (seen < n)]
        mask = np.ones(n, dtype=bool)
        mask[seen] = False
        cand = cand[mask]
    order = np.argsort(-scores[cand], kind="stable")[:k]
    sel = cand[order]
    return sel + item_base, scores[sel]


def merge_topk(lists: Sequence[Tuple[np.ndarray, np.ndarray]], k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k of the union of per-shard top-k lists; ties -> lower global index
    (shards are contiguous index ranges)."""
    idx = np.concatenate([l[0] for l in lists])
    sc = np.concatenate([l[1] for l in lists])
    order = np.lexsort((idx, -sc))[:k]
    return idx[order], sc[order]


# --------------------------------------------------------------------------
# metrics
# --------------------------------------------------------------------------
def ndcg_tasks(recommended: Sequence, relevant: Set, k: int) -> float:
    """src/evaluation/tasks.py:718-747 (IDCG over min(len(relevant), k))."""
    if not relevant:
        return 0.0
    dcg = 0.0
    for i, item in enumerate(recommended[:k], 1):
        if item in relevant:
            dcg += 1.0 / np.log2(i + 1)
    num_relevant = min(len(relevant), k)
    idcg = sum(1.0 / np.log2(i + 2) for i in range(num_relevant))
    return dcg / idcg if idcg > 0 else 0.0


def retrieval_metrics(all_recs: Sequence[Sequence], all_pos: Sequence[Set], k: int) -> Dict[str, float]:
    """Accuracy block of TopKRetrievalEvaluator.evaluate
    (src/evaluation/tasks.py:567-635): precision denominator is len(recs),
    users without positives contribute zeros but stay in the mean."""
    n = len(all_recs)
    hits = np.zeros(n)
    pden = np.array([len(r) for r in all_recs], dtype=np.float32)
    rden = np.array([len(p) for p in all_pos], dtype=np.float32)
    mrr = np.zeros(n)
    ndcg = np.zeros(n)
    for i in range(n):
        rec_set, pos_set = set(all_recs[i]), set(all_pos[i])
        if not pos_set:
            continue
        hits[i] = len(rec_set & pos_set)
        for j, item in enumerate(all_recs[i], 1):
            if item in pos_set:
                mrr[i] = 1.0 / j
                break
        ndcg[i] = ndcg_tasks(list(all_recs[i]), pos_set, k)
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = hits / pden
        recall = hits / rden
    precision[np.isnan(precision)] = 0.0
    recall[np.isnan(recall)] = 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        f1 = 2 * precision * recall / (precision + recall)
    f1[np.isnan(f1)] = 0.0
    hit_rate = (hits > 0).astype(float)
    return {
        "avg_precision_at_k": float(np.mean(precision)) if n else 0.0,
        "avg_recall_at_k": float(np.mean(recall)) if n else 0.0,
        "avg_f1_at_k": float(np.mean(f1)) if n else 0.0,
        "avg_hit_rate_at_k": float(np.mean(hit_rate)) if n else 0.0,
        "avg_ndcg_at_k": float(np.mean(ndcg)) if n else 0.0,
        "avg_mrr": float(np.mean(mrr)) if n else 0.0,
        "num_users_evaluated": n,
    }


def precision_at_k(recommended, relevant, k):
    """src/evaluation/metrics.py:11-35 (denominator k)."""
    if not recommended or k == 0:
        return 0.0
    return sum(1 for it in recommended[:k] if it in relevant) / k


def recall_at_k(recommended, relevant, k):
    """src/evaluation/metrics.py:37-61."""
    if not relevant or k == 0:
        return 0.0
    return sum(1 for it in recommended[:k] if it in relevant) / len(relevant)


def ndcg_metrics(recommended, relevant, k):
    """src/evaluation/metrics.py:63-100 (IDCG over the sorted hit vector)."""
    rel = [1 if it in relevant else 0 for it in recommended[:k]]
    if sum(rel) == 0:
        return 0.0
    dcg = lambda s: sum(v / np.log2(i + 2) for i, v in enumerate(s))
    return dcg(rel) / dcg(sorted(rel, reverse=True))


def average_precision(recommended, relevant):
    """src/evaluation/metrics.py:102-133."""
    if not relevant:
        return 0.0
    precisions, hits = [], 0
    for i, it in enumerate(recommended):
        if it in relevant:
            hits += 1
            precisions.append(hits / (i + 1))
    return sum(precisions) / len(relevant) if precisions else 0.0


def mrr(recommendations, relevant_items):
    """src/evaluation/advanced_metrics.py:15-45."""
    rr = []
    for recs, rel in zip(recommendations, relevant_items):
        for i, it in enumerate(recs):
            if it in rel:
                rr.append(1.0 / (i + 1))
                break
        else:
            rr.append(0.0)
    return float(np.mean(rr)) if rr else 0.0


def hit_rate(recommendations, relevant_items):
    """src/evaluation/advanced_metrics.py:47-69."""
    if not recommendations:
        return 0.0
    return sum(1 for recs, rel in zip(recommendations, relevant_items) if any(it in rel for it in recs)) / len(recommendations)


# --------------------------------------------------------------------------
# split / folded algebra used by the kernels, restated
# in fp64 so tests can check the algebra independently of bf16 rounding
# --------------------------------------------------------------------------
def fold_batchnorm(sd, use_batch_norm: bool, dt=np.float64):
    """BN sits after the activation, so it folds into the *next* Linear:
    s = g/sqrt(var+eps), t = b - mean*s, W' = W diag(s), b' = b + W t."""
    hidden, last = prediction_layout(sd, use_batch_norm)
    ws, bs = [], []
    scale, shift = None, None
    for i in hidden + [last]:
        p = f"prediction_network.{i}"
        w = sd[p + ".weight"].astype(dt)
        b = sd[p + ".bias"].astype(dt)
        if scale is not None:
            b = b + w @ shift
            w = w * scale[None, :]
        ws.append(w)
        bs.append(b)
        if use_batch_norm and i != last:
            bn = f"prediction_network.{i + 2}"
            scale = sd[bn + ".weight"].astype(dt) / np.sqrt(sd[bn + ".running_var"].astype(dt) + 1e-5)
            shift = sd[bn + ".bias"].astype(dt) - sd[bn + ".running_mean"].astype(dt) * scale
        else:
            scale, shift = None, None
    return ws, bs


# --------------------------------------------------------------------------
# reduced-precision emulation of the tcgen05 kernel's rounding points
# --------------------------------------------------------------------------
def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16 (8-bit significand), returned as float64."""
    a = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    a = (a + 0x7FFF + ((a >> 16) & 1)) & 0xFFFF0000
    return a.astype(np.uint32).view(np.float32).astype(np.float64)


def round_fp16(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)


def forward_pairs_lowp(sd, cfg, user_idx, item_idx, tag_idx, vis=None, txt=None, num=None, rnd=round_bf16,
                       return_logit=False):
    """The same forward with the operand roundings of the fused tcgen05 kernel
    (csrc/score_tc.cu): BatchNorm folded into the next Linear, the fused vector
    (gated) or layer-1 activation, every hidden activation and the hidden-layer
    weights rounded to the 16-bit operand format, fp32/fp64 accumulation,
    fp32 biases, last Linear in full precision.  Used to separate "the kernel
    computes what it documents" (kernel vs this, tight) from "what 16-bit
    operands cost" (this vs forward_pairs, reported)."""
    dt = np.float64
    act = cfg.get("fusion_activation", "relu")
    feats = modality_features(sd, act, np.asarray(user_idx), np.asarray(item_idx), np.asarray(tag_idx), vis, txt, num, dt)
    ft = cfg.get("fusion_type", "concatenate")
    ws, bs = fold_batchnorm(sd, bool(cfg.get("use_batch_norm", True)))
    if ft == "gated" and feats[0].shape[1] != 64:
        # embedding_dim != 64 (F_GATEDW): layer 1 is linear in the fused vector and the gate weights sum to 1, so it is the
        # gate-weighted sum of one per-user partial (fp32) and M - 1 per-item partials W1 f_m + b1 (stored in 16 bit);
        # the sum goes through the activation and is rounded once as the layer-2 operand
        # The item partials are stored in fp16 whatever the operand format, and their gate-weighted sum runs on packed
        # fp16 FMAs (one rounding per multiply / fused multiply-add, modality order); the user term stays fp32.
        g = gated_fusion(sd, feats, dt, return_gates=True)
        acc = None
        for m in range(1, len(feats)):
            gm = round_fp16(g[:, m:m + 1])
            qm = round_fp16(np.clip(feats[m] @ ws[0].T + bs[0], -65504.0, 65504.0))
            acc = round_fp16(gm * qm) if acc is None else round_fp16(gm * qm + acc)
        z1 = g[:, :1] * (feats[0] @ ws[0].T + bs[0]) + acc
        h = rnd(activation(z1, act))
        start = 1
    elif ft == "gated":
        x = rnd(gated_fusion(sd, feats, dt))
        h = x
        start = 0
    elif ft == "concatenate":
        # layer 1 is split into a per-user partial (fp32) and a per-item partial (+ bias, stored in 16 bit),
        # SURVEY.md A3; their sum goes through the activation and is rounded once more as the layer-2 operand
        D = feats[0].shape[1]
        x = np.concatenate(feats, axis=1)
        pu = x[:, :D] @ ws[0][:, :D].T
        pi = rnd(x[:, D:] @ ws[0][:, D:].T + bs[0])
        h = rnd(activation(pu + pi, act))
        start = 1
    elif ft == "attention":
        # The kernel rounds acc = sum over tokens of the normalised rows to the operand format and
        # folds the LayerNorm affine and the mean into layer 1: W1' = W1 diag(ln_w / M) (rounded), b1' = b1 + W1 ln_b.
        M = len(feats)
        g = sd["fusion_layer.norm.weight"].astype(dt); beta = sd["fusion_layer.norm.bias"].astype(dt)
        # (attention_token_sum_split restates the kernel's split evaluation order; with exact storage it equals this)
        # the token sum itself comes from the front end's 16-bit register MMAs: attention_token_sum_mma restates them
        h = rnd(attention_token_sum_mma(sd, feats, int(cfg.get("num_attention_heads", 4)), rnd, dt))
        ws = list(ws); bs = list(bs)
        bs[0] = bs[0] + ws[0] @ beta
        ws[0] = ws[0] * (g / M)[None, :]
        start = 0
    else:
        raise ValueError("no reduced-precision kernel for fusion type " + ft)
    for li in range(start, len(ws) - 1):
        z = h @ rnd(ws[li]).T + bs[li]
        a = activation(z, act)
        h = rnd(a) if li < len(ws) - 2 else a
    z = (h @ ws[-1].T + bs[-1])[:, 0]
    if return_logit:
        return z
    return final_activation(z, cfg.get("final_activation", "sigmoid"))


# --------------------------------------------------------------------------
# sampled evaluation protocol: candidate construction
# --------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def mix64(z: int) -> int:
    """splitmix64 step (Python ints, 64-bit wrap-around)."""
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _feistel_perm(ku: int, n_items: int):
    bits = 1
    while (1 << bits) < n_items:
        bits += 1
    h = (bits + 1) >> 1
    mask = (1 << h) - 1
    rk = [mix64((ku + r) & _M64) >> 32 for r in range(4)]

    def at(j: int) -> int:
        x = j
        while True:
            L, R = x >> h, x & mask
            for r in range(4):
                f = mix64((rk[r] << 32) | R) & 0xFFFFFFFF & mask
                L, R = R, L ^ f
            x = (L << h) | R
            if x < n_items:
                return x
    return at


def sample_candidates(global_user: int, positives: Sequence[int], n_items: int, n_neg: int, seed: int,
                      stride: int = 1024) -> List[int]:
    """Candidate list of one user for the sampled protocol (reference
    src/evaluation/tasks.py:181-224 'random' strategy + the shuffle of :336-342),
    with the reference's unreproducible salted-hash seeding replaced by a pure
    function of (seed, global user index): negatives = the first n_neg images of a
    keyed Feistel permutation of range(n_items) that are not positives (uniform,
    without replacement); final order = ascending 64-bit hash of (user key, item),
    ties -> lower item.  Pure-Python restatement of csrc/sampling.cu."""
    pos = sorted(int(p) for p in positives)[:stride]
    posset = set(int(p) for p in positives)
    ku = mix64((seed & _M64) ^ mix64(global_user & _M64))
    want = max(0, min(n_neg, stride - len(pos), n_items - len(posset)))
    at = _feistel_perm(ku, n_items)
    negs, j = [], 0
    while len(negs) < want:
        x = at(j)
        j += 1
        if x not in posset:
            negs.append(x)
    cand = pos + negs
    key = lambda it: (mix64(ku ^ 0xD1B54A32D192ED03 ^ ((it & 0xFFFFFFFF) << 1)), it)
    return sorted(cand, key=key)


def sampling_weights(test_items: Sequence[int], n_items: int, strategy: str) -> np.ndarray:
    """Per-item sampling weight of the 'popularity' / 'popularity_inverse' strategies
    (src/evaluation/tasks.py:227-243, 266-280): the number of rows of the TEST table that hold the item (default 1
    for items that never occur there), or its reciprocal."""
    cnt = np.bincount(np.asarray(test_items, dtype=np.int64), minlength=n_items).astype(np.float64)
    cnt[cnt == 0] = 1.0
    return cnt if strategy == "popularity" else 1.0 / cnt


def sample_candidates_weighted(global_user: int, positives: Sequence[int], weights: np.ndarray, n_neg: int, seed: int,
                               stride: int = 1024) -> List[int]:
    """Candidate list of one user for the popularity-biased strategies (src/evaluation/tasks.py:225-308: negatives drawn
    without replacement with probability proportional to the weight, ``np.random.choice(..., replace=False, p=...)``),
    made reproducible like ``sample_candidates``: item i gets the exponential key log(u_i) / w_i with u_i a hash-uniform
    of (seed, user, item); the n_neg largest keys among the non-positive items are the sample (Efraimidis-Spirakis:
    the same distribution as successive weighted draws without replacement).  Ties -> lower item.  Final order as in
    ``sample_candidates``.  Pure-Python restatement of csrc/sampling.cu::weighted_candidates_kernel."""
    n_items = len(weights)
    pos = sorted(int(p) for p in positives)[:stride]
    posset = set(int(p) for p in positives)
    ku = mix64((seed & _M64) ^ mix64(global_user & _M64))
    want = max(0, min(n_neg, stride - len(pos), n_items - len(posset)))
    keyed = []
    for it in range(n_items):
        if it in posset:
            continue
        u = ((mix64(ku ^ 0xA0761D6478BD642F ^ ((it & 0xFFFFFFFF) << 1)) >> 11) + 0.5) * (1.0 / 9007199254740992.0)
        keyed.append((-(math.log(u) / float(weights[it])), it))
    keyed.sort()
    negs = [it for _, it in keyed[:want]]
    cand = pos + negs
    key = lambda it: (mix64(ku ^ 0xD1B54A32D192ED03 ^ ((it & 0xFFFFFFFF) << 1)), it)
    return sorted(cand, key=key)


def rank_candidates(scores: Sequence[float], candidates: Sequence[int], k: int) -> List[int]:
    """Stable descending sort over the candidate order, first k
    (src/inference/recommender.py:97-106 with candidates=...)."""
    order = sorted(range(len(candidates)), key=lambda i: -scores[i])      # Python's sort is stable
    return [int(candidates[i]) for i in order[:k]]


# --------------------------------------------------------------------------
# beyond-accuracy metrics
# --------------------------------------------------------------------------
def novelty_tables(interaction_items: Sequence[int], interaction_users: Sequence[int], n_items: int):
    """Per-item tables of NoveltyMetrics (src/evaluation/novelty.py:26-65): popularity = interaction count per item
    (``value_counts`` at tasks.py:646), self-information -log2(max(pop/total, 1e-10)) (:149-178), IIF
    log(n_users / (count + 1e-10)) (:180-206); NaN where the item never occurs in the interactions."""
    cnt = np.bincount(np.asarray(interaction_items, dtype=np.int64), minlength=n_items).astype(np.float64)
    total = cnt.sum()
    n_users = len(set(int(u) for u in interaction_users))
    si = np.full(n_items, np.nan)
    iif = np.full(n_items, np.nan)
    has = cnt > 0
    if total > 0:
        si[has] = -np.log2(np.maximum(cnt[has] / total, 1e-10))
    if n_users > 0:
        iif[has] = np.log(n_users / (cnt[has] + 1e-10))
    return si, iif, int(has.sum())


def novelty_metrics(rec_lists: Sequence[Sequence[int]], histories: Sequence[Set[int]], si, iif, n_pop: int) -> Dict[str, float]:
    """Aggregation of TopKRetrievalEvaluator.evaluate (src/evaluation/tasks.py:674-714) over per-user
    NoveltyMetrics.calculate_metrics (novelty.py:84-147) and _calculate_personalization (tasks.py:402-427),
    written as the direct loops (personalization as the explicit mean over all user pairs)."""
    a_si, a_iif, a_cov, a_pn = [], [], [], []
    for recs, hist in zip(rec_lists, histories):
        recs = [int(r) for r in recs if r >= 0]
        if not recs:
            continue
        v = [si[r] for r in recs if not np.isnan(si[r])]
        a_si.append(np.mean(v) if v else 0.0)
        v = [iif[r] for r in recs if not np.isnan(iif[r])]
        a_iif.append(np.mean(v) if v else 0.0)
        a_cov.append(len(set(recs)) / n_pop if n_pop else 0.0)
        a_pn.append(len([r for r in recs if r not in hist]) / len(recs))
    n = len(rec_lists)
    if n <= 1:
        pers = 1.0 if n == 1 else 0.0
    else:
        sets = [set(int(r) for r in recs if r >= 0) for recs in rec_lists]
        tot = 0.0
        for a in range(n):
            for b in range(a + 1, n):
                if sets[a] and sets[b]:
                    tot += len(sets[a] & sets[b]) / math.sqrt(len(sets[a]) * len(sets[b]))
        pers = 1.0 - tot / (n * (n - 1) / 2)
    m = lambda x: float(np.mean(x)) if x else 0.0
    return {"avg_self_information": m(a_si), "avg_iif": m(a_iif), "avg_catalog_coverage": m(a_cov),
            "avg_personalization": pers, "avg_personalized_novelty": m(a_pn)}


# --------------------------------------------------------------------------
# ranking task (--eval_task ranking)
# --------------------------------------------------------------------------
def gini_coefficient(counts: Sequence[int]) -> float:
    """AdvancedMetrics.calculate_gini_coefficient (src/evaluation/advanced_metrics.py:72-105) of the dict values."""
    c = np.sort(np.asarray(list(counts), dtype=np.float64))
    n = len(c)
    if n == 0 or c.sum() == 0:
        return 0.0
    idx = np.arange(1, n + 1)
    return float((2 * np.sum(idx * c)) / (n * c.sum()) - (n + 1) / n)


def intra_list_similarity(items: Sequence[int], emb: np.ndarray) -> float:
    """NoveltyMetrics.calculate_diversity (src/evaluation/novelty.py:295-340): mean of the upper triangle of the
    cosine-similarity matrix of the listed items' embeddings; < 2 embeddings -> 0.0 (zero rows count as missing)."""
    rows = [np.asarray(emb[i], dtype=np.float64) for i in items if i >= 0 and np.any(emb[i] != 0)]
    if len(rows) < 2:
        return 0.0
    X = np.stack(rows)
    X = X / np.linalg.norm(X, axis=1, keepdims=True)
    S = X @ X.T
    iu = np.triu_indices(len(rows), k=1)
    return float(S[iu].mean())


def ranking_task_metrics(users: Sequence[str], test_items: Sequence[Sequence[str]],
                         score_fn, k: int) -> Dict[str, object]:
    """TopKRankingEvaluator.evaluate (src/evaluation/tasks.py:776-901) as the direct per-user loop: the user's test
    items are scored one by one (``get_item_score``), sorted by score with Python's stable descending sort
    (:830), and every test item counts as relevant -- ranks are 1..n (:835-837), MRR is 1/ranks[0] (:845),
    hit rate counts ranks <= k over len(test_items) (:848-849), NDCG is ``ndcg_tasks`` of the sorted list against
    the SET of test items (:853-854; duplicates in the table shrink the ideal, not the gain).  ``users`` are in
    ``groupby('user_id')`` order; ``score_fn(user, item) -> float``.  Means / population std as :877-893."""
    per = {"avg_rank": [], "median_rank": [], "mrr": [], "hit_rate_at_k": [], "ndcg_at_k": []}
    predictions = {}
    for u, items in zip(users, test_items):
        scored = [(str(i), float(score_fn(str(u), str(i)))) for i in items]
        if not scored:
            for v in per.values():
                v.append(0.0)
            continue
        scored.sort(key=lambda x: x[1], reverse=True)
        predictions[str(u)] = scored
        ranks = list(range(1, len(scored) + 1))
        per["avg_rank"].append(float(np.mean(ranks)))
        per["median_rank"].append(float(np.median(ranks)))
        per["mrr"].append(1.0 / ranks[0])
        per["hit_rate_at_k"].append(sum(1 for r in ranks if r <= k) / len(items))
        per["ndcg_at_k"].append(ndcg_tasks([i for i, _ in scored], set(str(i) for i in items), k))
    out: Dict[str, object] = {}
    for name, vals in per.items():
        out[f"avg_{name}"] = float(np.mean(vals)) if vals else 0.0
        out[f"std_{name}"] = float(np.std(vals)) if vals else 0.0
    out["num_users_evaluated"] = len(users)
    out["predictions"] = predictions
    return out


# --------------------------------------------------------------------------
# K4 register merge: the compare-exchange network, lane by lane
# --------------------------------------------------------------------------
def merge_top64_network(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Restatement of ``merge_top64`` in csrc/simt_kernels.cu (the k <= 64 path of ``pxr_merge_topk``; the reference
    has no counterpart -- it sorts one full score list per user, src/inference/recommender.py:105).  ``a`` and ``b`` are
    descending lists of 64 distinct uint64 keys (0 = padding); slot j lives in lane j & 31, register j >> 5.  Returns
    the best 64 keys of the union, descending: max(A[j], B[63 - j]) is bitonic, six compare-exchange stages sort it."""
    lane = np.arange(32)
    a0, a1, b0, b1 = a[:32].copy(), a[32:].copy(), b[:32], b[32:]
    r0, r1 = b1[lane ^ 31], b0[lane ^ 31]                     # reversed B: one lane-mirror shuffle, registers swapped
    x0, x1 = np.maximum(a0, r0), np.maximum(a1, r1)
    x0, x1 = np.maximum(x0, x1), np.minimum(x0, x1)           # distance 32: inside the lane
    d = 16
    while d >= 1:                                             # distances 16 .. 1: shfl.xor partners
        p0, p1 = x0[lane ^ d], x1[lane ^ d]
        keep_max = (lane & d) == 0
        take0, take1 = (p0 > x0) == keep_max, (p1 > x1) == keep_max
        x0, x1 = np.where(take0, p0, x0), np.where(take1, p1, x1)
        d >>= 1
    return np.concatenate([x0, x1])

"""PixelRec-shaped synthetic workload generator.

Two generators live here:

* ``det_*``: a counter-based, version-independent generator (SplitMix64 over
  uint64 numpy arithmetic).  Given (seed, stream name, shape) it always returns
  the same numbers on every numpy / platform, so the golden fixtures under
  ``tests/golden`` only need to store *outputs* of the reference, not weights.
* ``torch_*`` helpers used by ``bench.py`` to materialise the big
  Pixel200K/1M/8M-shaped tables directly on the GPU.

Weights are "trained-like" rather than the reference's ``xavier_uniform`` on
the embedding tables (reference ``src/models/multimodal.py:185-188``): with the
reference init every score collapses into [0.47, 0.52] and a top-K parity test
would be vacuous (SURVEY.md §7 "hard parts").
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

SEED = 20261018

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        z = z ^ (z >> np.uint64(31))
    return z


def _stream_base(seed: int, name: str) -> np.uint64:
    h = zlib.crc32(name.encode("utf-8")) & 0xFFFFFFFF
    base = np.array([(seed & 0xFFFFFFFF) << 32 | h], dtype=np.uint64)
    return _splitmix64(base)[0]


def det_uniform(seed: int, name: str, shape, offset: int = 0) -> np.ndarray:
    """float64 uniforms in [0, 1), one per element, addressed by flat index."""
    n = int(np.prod(shape)) if len(tuple(np.atleast_1d(shape))) else 1
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        bits = _splitmix64(idx * np.uint64(0xD1342543DE82EF95) + _stream_base(seed, name))
    u = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return u.reshape(shape)


def det_normal(seed: int, name: str, shape) -> np.ndarray:
    """float64 standard normals via Box-Muller on two det_uniform streams."""
    u1 = det_uniform(seed, name + "/u1", shape)
    u2 = det_uniform(seed, name + "/u2", shape)
    u1 = np.maximum(u1, 1e-300)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def det_randint(seed: int, name: str, shape, high: int) -> np.ndarray:
    return np.minimum((det_uniform(seed, name, shape) * high).astype(np.int64), high - 1)


@dataclass
class ModelSpec:
    """Mirror of the reference ctor arguments that shape the weights
    (reference ``src/models/multimodal.py:42-66``)."""

    n_users: int
    n_items: int
    n_tags: int = 30
    num_numerical_features: int = 7
    embedding_dim: int = 64
    vision_dim: int = 512          # 0 = no vision modality
    language_dim: int = 384        # 0 = no language modality
    fusion_type: str = "concatenate"
    fusion_hidden_dims: List[int] = field(default_factory=lambda: [512, 256, 128])
    fusion_activation: str = "relu"
    use_batch_norm: bool = True
    projection_hidden_dim: Optional[int] = None
    final_activation: str = "sigmoid"
    num_attention_heads: int = 4

    @property
    def num_modalities(self) -> int:
        return 3 + (self.vision_dim > 0) + (self.language_dim > 0) + (self.num_numerical_features > 0)


def _linear(seed, name, out_f, in_f, sd, key):
    bound = 1.0 / np.sqrt(in_f)
    sd[key + ".weight"] = ((det_uniform(seed, name + ".w", (out_f, in_f)) * 2 - 1) * bound).astype(np.float32)
    sd[key + ".bias"] = ((det_uniform(seed, name + ".b", (out_f,)) * 2 - 1) * bound).astype(np.float32)


def make_state_dict(spec: ModelSpec, seed: int = SEED, logit_scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Random "trained-like" weights under the reference state_dict key names
    (SURVEY.md §8(a) A1)."""
    D = spec.embedding_dim
    sd: Dict[str, np.ndarray] = {}
    s = 1.0 / np.sqrt(D)
    sd["user_embedding.weight"] = (det_normal(seed, "user_embedding", (spec.n_users, D)) * s).astype(np.float32)
    sd["item_embedding.weight"] = (det_normal(seed, "item_embedding", (spec.n_items, D)) * s).astype(np.float32)
    sd["tag_embedding.weight"] = (det_normal(seed, "tag_embedding", (spec.n_tags, D)) * s).astype(np.float32)

    def proj(prefix, in_dim):
        if spec.projection_hidden_dim:
            P = spec.projection_hidden_dim
            _linear(seed, prefix + ".0", P, in_dim, sd, prefix + ".0")
            _linear(seed, prefix + ".3", D, P, sd, prefix + ".3")
        else:
            _linear(seed, prefix + ".0", D, in_dim, sd, prefix + ".0")

    if spec.vision_dim > 0:
        proj("vision_projection", spec.vision_dim)
    if spec.language_dim > 0:
        proj("language_projection", spec.language_dim)
    if spec.num_numerical_features > 0:
        proj("numerical_projection", spec.num_numerical_features)

    M = spec.num_modalities
    if spec.fusion_type == "concatenate":
        fusion_in = M * D
    elif spec.fusion_type == "gated":
        _linear(seed, "gate", M, M * D, sd, "fusion_layer.gating_network.0")
        # trained gates are far from uniform: widen the logits
        sd["fusion_layer.gating_network.0.weight"] *= np.float32(3.0)
        fusion_in = D
    elif spec.fusion_type == "attention":
        b = 1.0 / np.sqrt(D)
        sd["fusion_layer.attention.in_proj_weight"] = (
            (det_uniform(seed, "attn.in_w", (3 * D, D)) * 2 - 1) * b * 1.5).astype(np.float32)
        sd["fusion_layer.attention.in_proj_bias"] = (
            (det_uniform(seed, "attn.in_b", (3 * D,)) * 2 - 1) * 0.1).astype(np.float32)
        sd["fusion_layer.attention.out_proj.weight"] = (
            (det_uniform(seed, "attn.out_w", (D, D)) * 2 - 1) * b).astype(np.float32)
        sd["fusion_layer.attention.out_proj.bias"] = (
            (det_uniform(seed, "attn.out_b", (D,)) * 2 - 1) * 0.1).astype(np.float32)
        sd["fusion_layer.norm.weight"] = (1.0 + 0.1 * det_normal(seed, "attn.ln_w", (D,))).astype(np.float32)
        sd["fusion_layer.norm.bias"] = (0.1 * det_normal(seed, "attn.ln_b", (D,))).astype(np.float32)
        fusion_in = D
    else:
        raise ValueError(f"Unknown fusion type: '{spec.fusion_type}'")

    stride = 4 if spec.use_batch_norm else 3
    in_dim = fusion_in
    for li, h in enumerate(spec.fusion_hidden_dims):
        base = li * stride
        _linear(seed, f"mlp{li}", h, in_dim, sd, f"prediction_network.{base}")
        if spec.use_batch_norm:
            bn = f"prediction_network.{base + 2}"
            sd[bn + ".weight"] = (1.0 + 0.1 * det_normal(seed, bn + ".g", (h,))).astype(np.float32)
            sd[bn + ".bias"] = (0.1 * det_normal(seed, bn + ".b", (h,))).astype(np.float32)
            sd[bn + ".running_mean"] = (0.2 * det_normal(seed, bn + ".m", (h,))).astype(np.float32)
            sd[bn + ".running_var"] = (0.5 + det_uniform(seed, bn + ".v", (h,))).astype(np.float32)
            sd[bn + ".num_batches_tracked"] = np.array(100, dtype=np.int64)
        in_dim = h
    last = len(spec.fusion_hidden_dims) * stride
    _linear(seed, "mlp_out", 1, in_dim, sd, f"prediction_network.{last}")
    sd[f"prediction_network.{last}.weight"] *= np.float32(logit_scale)
    return sd


def final_linear_key(spec: ModelSpec) -> str:
    stride = 4 if spec.use_batch_norm else 3
    return f"prediction_network.{len(spec.fusion_hidden_dims) * stride}"


def make_item_features(spec: ModelSpec, seed: int = SEED) -> Dict[str, np.ndarray]:
    """Per-item cached features as the hoisted frozen backbones would emit them
    (SURVEY.md §8(d)): CLIP-pooled-like vision vectors of norm 10, unit-norm
    SBERT-like text vectors, standardised numericals, Zipf tags."""
    NI = spec.n_items
    out: Dict[str, np.ndarray] = {}
    if spec.vision_dim > 0:
        v = det_normal(seed, "feat.vis", (NI, spec.vision_dim))
        out["vis"] = (10.0 * v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    if spec.language_dim > 0:
        t = det_normal(seed, "feat.txt", (NI, spec.language_dim))
        out["txt"] = (t / np.linalg.norm(t, axis=1, keepdims=True)).astype(np.float32)
    if spec.num_numerical_features > 0:
        out["num"] = det_normal(seed, "feat.num", (NI, spec.num_numerical_features)).astype(np.float32)
    ranks = np.arange(1, spec.n_tags + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** -1.2)
    cdf /= cdf[-1]
    out["tag_idx"] = np.searchsorted(cdf, det_uniform(seed, "feat.tag", (NI,))).astype(np.int64)
    return out


def make_histories(n_users: int, n_items: int, seed: int = SEED, mean_log: float = 2.6,
                   sigma_log: float = 0.6, lo: int = 5, hi: int = 500):
    """Leave-one-out histories (reference ``src/data/splitting.py:282-337``):
    per user a Zipf(1.0)-popular item set; the last item is the test positive,
    the second-last validation, the rest the train history used by
    ``filter_seen``.  Returns CSR (indptr int64, idx int32 sorted ascending) for
    train, plus ``test_item`` (int32, one per user)."""
    n = np.clip(np.exp(mean_log + sigma_log * det_normal(seed, "hist.n", (n_users,))), lo, min(hi, n_items // 2))
    n = n.astype(np.int64)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(n, out=indptr[1:])
    total = int(indptr[-1])
    ranks = np.arange(1, n_items + 1, dtype=np.float64)
    cdf = np.cumsum(1.0 / ranks)
    cdf /= cdf[-1]
    draws = np.searchsorted(cdf, det_uniform(seed, "hist.items", (total,))).astype(np.int64)
    perm = np.argsort(det_uniform(seed, "hist.perm", (n_items,)), kind="stable")
    draws = perm[np.minimum(draws, n_items - 1)]
    train_indptr = np.zeros(n_users + 1, dtype=np.int64)
    train_chunks, test_item = [], np.full(n_users, -1, dtype=np.int32)
    for u in range(n_users):
        items = draws[indptr[u]:indptr[u + 1]]
        _, first = np.unique(items, return_index=True)
        seq = items[np.sort(first)]                    # de-duplicated, in "time" order
        if len(seq) >= 3:
            test_item[u] = seq[-1]
            tr = np.sort(seq[:-2])
        else:
            tr = np.sort(seq)
        train_chunks.append(tr.astype(np.int32))
        train_indptr[u + 1] = train_indptr[u] + len(tr)
    train_idx = np.concatenate(train_chunks) if train_chunks else np.zeros(0, np.int32)
    return train_indptr, train_idx, test_item


def user_ids(n: int) -> List[str]:
    """Zero-padded so LabelEncoder's lexicographic order equals numeric order
    (reference ``src/data/dataset.py:142-148``)."""
    return [f"u{i:08d}" for i in range(n)]


def item_ids(n: int) -> List[str]:
    return [f"i{i:08d}" for i in range(n)]


def apply_logit_calibration(sd: Dict[str, np.ndarray], spec: ModelSpec, mean: float, std: float,
                            target_std: float = 2.0) -> None:
    """Rescale the output Linear in place so that pre-activation logits measured
    by the caller as (mean, std) become (0, target_std): a trained model spreads
    its scores over (0, 1) instead of the random-init [0.47, 0.52] band
    ("bf16 tolerance vs. score compression")."""
    key = final_linear_key(spec)
    scale = np.float32(target_std / max(std, 1e-12))
    sd[key + ".weight"] = (sd[key + ".weight"] * scale).astype(np.float32)
    sd[key + ".bias"] = ((sd[key + ".bias"] - np.float32(mean)) * scale).astype(np.float32)


# --------------------------------------------------------------------------
# torch generators for the big Pixel200K/1M/8M-shaped workloads (bench.py)
# --------------------------------------------------------------------------
def torch_workload(spec: ModelSpec, device, seed: int = SEED, with_histories: bool = True):
    """Same distributions as the det_* generators but drawn
    with a seeded ``torch.Generator`` on ``device`` so 1M-8M-user tables take
    milliseconds.  Returns (state_dict of torch tensors, feature dict, history
    dict or None).  The small dense weights still come from ``make_state_dict``
    so they are identical to what the tests use."""
    import dataclasses
    import torch

    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    D = spec.embedding_dim
    small = dataclasses.replace(spec, n_users=1, n_items=1)
    sd = {k: torch.from_numpy(np.asarray(v)).to(dev) for k, v in make_state_dict(small, seed=seed).items()}
    s = 1.0 / np.sqrt(D)
    sd["user_embedding.weight"] = torch.randn((spec.n_users, D), generator=g, device=dev) * s
    sd["item_embedding.weight"] = torch.randn((spec.n_items, D), generator=g, device=dev) * s
    NI = spec.n_items
    feats = {}
    if spec.vision_dim > 0:
        v = torch.randn((NI, spec.vision_dim), generator=g, device=dev)
        feats["vis"] = 10.0 * v / v.norm(dim=1, keepdim=True)
    if spec.language_dim > 0:
        t = torch.randn((NI, spec.language_dim), generator=g, device=dev)
        feats["txt"] = t / t.norm(dim=1, keepdim=True)
    if spec.num_numerical_features > 0:
        feats["num"] = torch.randn((NI, spec.num_numerical_features), generator=g, device=dev)
    ranks = torch.arange(1, spec.n_tags + 1, device=dev, dtype=torch.float64)
    cdf = torch.cumsum(ranks ** -1.2, 0)
    cdf = cdf / cdf[-1]
    feats["tag_idx"] = torch.searchsorted(cdf, torch.rand(NI, generator=g, device=dev, dtype=torch.float64)).clamp_(max=spec.n_tags - 1)
    hist = None
    if with_histories:
        hist = torch_histories(spec.n_users, NI, g, dev)
    return sd, feats, hist


def torch_histories(n_users: int, n_items: int, g, dev, mean_log: float = 2.6, sigma_log: float = 0.6,
                    lo: int = 5, hi: int = 500):
    """Vectorised leave-one-out histories: per user n_u ~ clip(LogNormal) Zipf(1.0)
    draws, de-duplicated; one extra draw is the held-out test positive and is
    removed from the train history.  Train CSR is ascending inside each user."""
    import torch
    n = torch.exp(mean_log + sigma_log * torch.randn(n_users, generator=g, device=dev)).clamp_(lo, min(hi, max(lo, n_items // 2))).long()
    owner = torch.repeat_interleave(torch.arange(n_users, device=dev), n)
    cdf = torch.cumsum(1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float64), 0)
    cdf = cdf / cdf[-1]
    perm = torch.randperm(n_items, generator=g, device=dev)
    draw = lambda m: perm[torch.searchsorted(cdf, torch.rand(m, generator=g, device=dev, dtype=torch.float64)).clamp_(max=n_items - 1)]
    items = draw(int(owner.numel()))
    test_item = draw(n_users)
    key = owner * n_items + items
    key = key[key != torch.arange(n_users, device=dev)[owner] * n_items + test_item[owner]]
    key = torch.unique(key)                                    # sorted: by user, then item
    u, it = key // n_items, key % n_items
    counts = torch.bincount(u, minlength=n_users)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(counts, 0)
    return dict(train_indptr=indptr, train_idx=it.to(torch.int32), test_item=test_item.to(torch.int32))


# --------------------------------------------------------------------------
# "trained-like" conditioning of a random model (workload generation only)
# --------------------------------------------------------------------------
def condition_like_trained(sd, spec: ModelSpec, feats, n_users: int = 64, n_items: int = 512, target_std: float = 2.0):
    """Make a random-init model numerically resemble a trained one, in place.

    Two things a trained checkpoint has and a random one lacks:
      * BatchNorm running statistics that MATCH the activations they normalise
        (training sets them to the batch statistics).  Random running stats leave a
        large common offset in every layer, so the pair-dependent part of the logit
        is a small difference of large numbers and any reduced-precision path looks
        ~10x worse than it would on a real model.
      * logits spread over several units (scores over (0, 1)) instead of the
        random-init [0.47, 0.52] band.
    ``sd`` / ``feats`` are dicts of torch tensors (any device) or numpy arrays (then
    converted and written back as numpy).  Statistics come from a sample of
    ``n_users x n_items`` pairs pushed through plain torch ops: this is workload
    generation, not the scoring path."""
    import torch
    import torch.nn.functional as F

    as_np = not isinstance(next(iter(sd.values())), torch.Tensor)
    T = {k: torch.as_tensor(np.asarray(v) if as_np else v) for k, v in sd.items()}
    Ft = {k: torch.as_tensor(np.asarray(v) if as_np else v) for k, v in feats.items()}
    dev = T["user_embedding.weight"].device
    nu, ni = min(n_users, spec.n_users), min(n_items, spec.n_items)
    uu = torch.arange(nu, device=dev).repeat_interleave(ni)
    ii = torch.arange(ni, device=dev).repeat(nu)
    act = {"relu": F.relu, "gelu": F.gelu, "tanh": torch.tanh, "leaky_relu": lambda x: F.leaky_relu(x, 0.01),
           "silu": F.silu}.get((spec.fusion_activation or "relu").lower(), F.relu)

    def proj(x, prefix):
        y = act(F.linear(x, T[prefix + ".0.weight"].double(), T[prefix + ".0.bias"].double()))
        if prefix + ".3.weight" in T:
            y = act(F.linear(y, T[prefix + ".3.weight"].double(), T[prefix + ".3.bias"].double()))
        return y

    toks = [T["user_embedding.weight"].double()[uu], T["item_embedding.weight"].double()[ii],
            T["tag_embedding.weight"].double()[Ft["tag_idx"][ii].long()]]
    for key, prefix in (("vis", "vision_projection"), ("txt", "language_projection"), ("num", "numerical_projection")):
        if key in Ft and prefix + ".0.weight" in T:
            toks.append(proj(Ft[key][ii].double(), prefix))
    if spec.fusion_type == "concatenate":
        x = torch.cat(toks, 1)
    elif spec.fusion_type == "gated":
        g = torch.softmax(F.linear(torch.cat(toks, 1), T["fusion_layer.gating_network.0.weight"].double(),
                                   T["fusion_layer.gating_network.0.bias"].double()), -1)
        x = (torch.stack(toks, 1) * g.unsqueeze(-1)).sum(1)
    else:
        X = torch.stack(toks, 0)
        Dm = X.shape[-1]
        a, _ = F.multi_head_attention_forward(
            X, X, X, Dm, spec.num_attention_heads, T["fusion_layer.attention.in_proj_weight"].double(),
            T["fusion_layer.attention.in_proj_bias"].double(), None, None, False, 0.0,
            T["fusion_layer.attention.out_proj.weight"].double(), T["fusion_layer.attention.out_proj.bias"].double(),
            training=False, need_weights=False)
        x = F.layer_norm(X + a, (Dm,), T["fusion_layer.norm.weight"].double(), T["fusion_layer.norm.bias"].double(),
                         1e-5).mean(0)
    stride = 4 if spec.use_batch_norm else 3
    for li in range(len(spec.fusion_hidden_dims)):
        p = f"prediction_network.{li * stride}"
        x = act(F.linear(x, T[p + ".weight"].double(), T[p + ".bias"].double()))
        if spec.use_batch_norm:
            bn = f"prediction_network.{li * stride + 2}"
            mean, var = x.mean(0), x.var(0, unbiased=False) + 1e-3
            T[bn + ".running_mean"] = mean.to(T[bn + ".running_mean"].dtype)
            T[bn + ".running_var"] = var.to(T[bn + ".running_var"].dtype)
            x = (x - mean) / torch.sqrt(var + 1e-5) * T[bn + ".weight"].double() + T[bn + ".bias"].double()
    last = f"prediction_network.{len(spec.fusion_hidden_dims) * stride}"
    z = F.linear(x, T[last + ".weight"].double(), T[last + ".bias"].double())[:, 0]
    scale = target_std / max(float(z.std()), 1e-12)
    T[last + ".bias"] = ((T[last + ".bias"].double() - z.mean()) * scale).to(T[last + ".bias"].dtype)
    T[last + ".weight"] = (T[last + ".weight"].double() * scale).to(T[last + ".weight"].dtype)
    for k in list(sd.keys()):
        sd[k] = T[k].cpu().numpy() if as_np else T[k]
    return sd

"""CPU check of the algebra behind the fused attention front end (csrc/score_tc.cu: item_attn_kernel +
attn_user_setup + attn_tile): the per-item record / per-user constants / per-pair combination, restated in
numpy fp64 exactly as the kernels compute them, must reproduce the oracle's attention fusion
(reference src/models/layers.py:135-164) for every pair."""
import math

import numpy as np
import pytest

from oracle import pxr_oracle as orc
from pixelrec_multimodal_b200 import synthetic as syn


def _split_attention(sd, feats_list, heads):
    """feats_list: M arrays (B, D) in the reference's token order; token 0 is the user."""
    M = len(feats_list)
    B, D = feats_list[0].shape
    dh = D // heads
    w_in = sd["fusion_layer.attention.in_proj_weight"].astype(np.float64)
    b_in = sd["fusion_layer.attention.in_proj_bias"].astype(np.float64)
    w_o = sd["fusion_layer.attention.out_proj.weight"].astype(np.float64)
    b_o = sd["fusion_layer.attention.out_proj.bias"].astype(np.float64)
    ln_w = sd["fusion_layer.norm.weight"].astype(np.float64)
    ln_b = sd["fusion_layer.norm.bias"].astype(np.float64)
    scale = 1.0 / math.sqrt(dh)

    def qkv(x):
        y = x @ w_in.T + b_in
        return y[:, :D], y[:, D:2 * D], y[:, 2 * D:]

    def per_head_out(v):                       # U[h] = W_o[:, head h] v_h  -> (B, heads, D)
        return np.stack([v[:, h * dh:(h + 1) * dh] @ w_o[:, h * dh:(h + 1) * dh].T for h in range(heads)], axis=1)

    # ---- per-user constants (attn_user_setup)
    eu = feats_list[0]
    qu, ku, vu = qkv(eu)
    U0 = per_head_out(vu)
    C0 = eu + b_o
    S00 = np.stack([(qu[:, h * dh:(h + 1) * dh] * ku[:, h * dh:(h + 1) * dh]).sum(1) * scale for h in range(heads)], 1)
    # ---- per-item record (item_attn_kernel)
    nt = M - 1
    xs = feats_list[1:]
    q, k, v = zip(*[qkv(x) for x in xs])
    U = [per_head_out(vb) for vb in v]                              # nt x (B, heads, D)
    rec = []
    for a in range(nt):
        L = np.empty((B, heads)); Nbar = np.zeros((B, heads, D))
        for h in range(heads):
            sl = slice(h * dh, (h + 1) * dh)
            s = np.stack([(q[a][:, sl] * scale * k[b][:, sl]).sum(1) for b in range(nt)], 1)      # (B, nt)
            m = s.max(1, keepdims=True)
            e = np.exp(s - m)
            L[:, h] = (m + np.log(e.sum(1, keepdims=True)))[:, 0]
            pnorm = e / e.sum(1, keepdims=True)
            for b in range(nt):
                Nbar[:, h] += pnorm[:, b:b + 1] * U[b][:, h]
        C = xs[a] + b_o + Nbar.sum(1)
        rec.append(dict(C=C, Nbar=Nbar, U=U[a], q=q[a] * scale, k=k[a] * scale, L=L))
    # ---- per-pair combination (attn_tile)
    acc = np.zeros((B, D))
    S0 = np.stack([np.stack([(qu[:, h * dh:(h + 1) * dh] * rec[b]["k"][:, h * dh:(h + 1) * dh]).sum(1)
                             for h in range(heads)], 1) for b in range(nt)], 2)                       # (B, heads, nt)
    m = np.maximum(S00, S0.max(2))
    e0 = np.exp(S00 - m); eb = np.exp(S0 - m[:, :, None])
    inv = 1.0 / (e0 + eb.sum(2))
    y0 = C0 + ((e0 * inv)[:, :, None] * U0).sum(1)
    for b in range(nt):
        y0 = y0 + ((eb[:, :, b] * inv)[:, :, None] * rec[b]["U"]).sum(1)

    def ln_centered(y):
        t = y - y.mean(1, keepdims=True)
        return t / np.sqrt((t * t).mean(1, keepdims=True) + 1e-5)

    acc += ln_centered(y0)
    for a in range(nt):
        s_a0 = np.stack([(rec[a]["q"][:, h * dh:(h + 1) * dh] * ku[:, h * dh:(h + 1) * dh]).sum(1) for h in range(heads)], 1)
        w = 1.0 / (1.0 + np.exp(rec[a]["L"] - s_a0))
        y = rec[a]["C"] + (w[:, :, None] * (U0 - rec[a]["Nbar"])).sum(1)
        acc += ln_centered(y)
    return ln_b + ln_w / M * acc


@pytest.mark.parametrize("missing", [None, "num", "vis"])
def test_attention_split_matches_oracle(missing):
    kw = dict(n_users=40, n_items=60, fusion_type="attention")
    if missing == "num":
        kw["num_numerical_features"] = 0
    spec = syn.ModelSpec(**kw)
    sd = syn.make_state_dict(spec, seed=syn.SEED + 5)
    f = syn.make_item_features(spec, seed=syn.SEED + 5)
    rng = np.random.default_rng(3)
    u = rng.integers(0, spec.n_users, 500); i = rng.integers(0, spec.n_items, 500)
    vis = None if missing == "vis" else f["vis"][i]
    num = f["num"][i] if spec.num_numerical_features else None
    sdd = dict(sd)
    if missing == "vis":
        sdd = {k: v for k, v in sd.items() if not k.startswith("vision_projection")}
    feats = orc.modality_features(sdd, "relu", u, i, f["tag_idx"][i], vis, f["txt"][i], num)
    want = orc.attention_fusion(sdd, feats, spec.num_attention_heads, np.float64)
    got = _split_attention(sdd, feats, spec.num_attention_heads)
    assert len(feats) == (6 if missing is None else 5)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)

"""CPU check of the algebra behind the fused attention front end (csrc/score_tc.cu: item_attn_kernel +
attn_user_setup + attn_item_step): the per-item record / per-user operands / per-pair small GEMMs, restated in
numpy fp64 exactly as the kernels compute them (oracle.attention_token_sum_mma), must reproduce the oracle's
attention fusion (reference src/models/layers.py:135-164) for every pair; with the kernel's 16-bit operand
roundings switched on the result moves by less than the final 16-bit rounding of the fused vector itself."""
import math

import numpy as np
import pytest

from oracle import pxr_oracle as orc
from pixelrec_multimodal_b200 import synthetic as syn


def _split_attention(sd, feats_list, heads):
    """fused vector through the split form of the oracle (exact storage) + the LayerNorm affine and mean"""
    M = len(feats_list)
    acc = orc.attention_token_sum_mma(sd, feats_list, heads)
    return sd["fusion_layer.norm.bias"].astype(np.float64) + sd["fusion_layer.norm.weight"].astype(np.float64) / M * acc


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
    # token sum == the un-split oracle; the 16-bit MMA operands (scores, coefficient x value products, Wc, per-head
    # out-projected item values; Nc / xc carried as hi + lo) cost less than rounding the token sum itself once
    tok = orc.attention_fusion(sdd, feats, spec.num_attention_heads, np.float64, return_token_sum=True)
    np.testing.assert_allclose(orc.attention_token_sum_mma(sdd, feats, spec.num_attention_heads), tok, rtol=0, atol=1e-11)
    for rnd in (orc.round_bf16, orc.round_fp16):
        low = orc.attention_token_sum_mma(sdd, feats, spec.num_attention_heads, rnd)
        once = rnd(tok)
        assert 0 < np.sqrt(np.mean((low - tok) ** 2)) < 0.6 * np.sqrt(np.mean((once - tok) ** 2))
        assert np.abs(low - tok).max() < np.abs(once - tok).max()

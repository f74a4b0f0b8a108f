"""Beyond-accuracy metrics: pxr_novelty_metrics vs the direct-loop oracle restating
reference src/evaluation/novelty.py:84-226, 343-377 and tasks.py:402-427, 674-714."""
import numpy as np
import pytest

from oracle import pxr_oracle as orc


def _case(n_users, n_items, K, seed):
    rng = np.random.default_rng(seed)
    recs = np.stack([rng.permutation(n_items)[:K] for _ in range(n_users)]).astype(np.int32)
    recs[1, K // 2:] = -1                          # short list
    if n_users > 2:
        recs[2, :] = -1                            # empty list: no per-user metrics, still a row for personalization
    pop_items = rng.choice(n_items, int(n_items * 0.7), replace=False)          # 30 % of the catalogue never interacted with
    inter_u, inter_i = [], []
    for u in range(n_users):
        for it in rng.choice(pop_items, min(len(pop_items), int(rng.integers(0, 12))), replace=False):
            inter_u.append(u); inter_i.append(int(it))
    inter_u += [0, 0]; inter_i += [inter_i[0] if inter_i else 0] * 2            # duplicate interaction rows count twice
    return recs, np.array(inter_u), np.array(inter_i)


def test_oracle_personalization_identity():
    """sum_{u<v} cos(u, v) = (sum_i s_i^2 - #non-empty) / 2 with s_i = sum over lists containing i of 1/sqrt(|list|)."""
    recs, iu, ii = _case(40, 60, 7, 3)
    si, iif, n_pop = orc.novelty_tables(ii, iu, 60)
    want = orc.novelty_metrics(recs, [set() for _ in range(40)], si, iif, n_pop)["avg_personalization"]
    s = np.zeros(60); ne = 0
    for r in recs:
        r = set(int(x) for x in r if x >= 0)
        if r:
            ne += 1
            for x in r:
                s[x] += 1 / np.sqrt(len(r))
    got = 1 - ((s @ s - ne) / 2) / (40 * 39 / 2)
    assert abs(got - want) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("n_users,n_items,K", [(300, 500, 50), (33, 40, 10), (2, 9, 3), (1000, 5000, 64)])
def test_novelty_kernel_matches_oracle(n_users, n_items, K):
    import torch
    from pixelrec_multimodal_b200.evaluation import beyond_accuracy_metrics, novelty_tables
    recs, iu, ii = _case(n_users, n_items, K, n_users + K)
    hist = [set() for _ in range(n_users)]
    for u, it in zip(iu, ii):
        hist[int(u)].add(int(it))
    indptr = np.concatenate([[0], np.cumsum([len(h) for h in hist])]).astype(np.int64)
    idx = np.concatenate([np.array(sorted(h), dtype=np.int32) for h in hist] + [np.zeros(0, np.int32)])
    si, iif, n_pop = orc.novelty_tables(ii, iu, n_items)
    si2, iif2, n_pop2 = novelty_tables(ii, len(set(iu.tolist())), n_items)
    assert n_pop == n_pop2 and np.array_equal(si, si2, equal_nan=True) and np.array_equal(iif, iif2, equal_nan=True)
    want = orc.novelty_metrics(recs, hist, si, iif, n_pop)
    got = beyond_accuracy_metrics(torch.from_numpy(recs).cuda(), si, iif, n_pop, indptr, idx if len(idx) else np.zeros(1, np.int32))
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-9, (k, got[k], v)
    again = beyond_accuracy_metrics(torch.from_numpy(recs).cuda(), si, iif, n_pop, indptr, idx if len(idx) else np.zeros(1, np.int32))
    assert again == got                            # fixed-point accumulation: bitwise reproducible


def test_oracle_gini_known_answers():
    """reference tests/unit/src/evaluation/test_advanced_metrics.py:63-81"""
    assert orc.gini_coefficient([1, 1, 1, 1]) == pytest.approx(0.0)
    assert orc.gini_coefficient([4, 0, 0, 0]) == pytest.approx(1.0 - 1.0 / 4.0)
    assert 0 < orc.gini_coefficient([10, 5, 1, 1]) < 1
    assert orc.gini_coefficient([]) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("n_users,n_items,K", [(300, 120, 10), (5000, 700, 50), (3, 4, 2)])
def test_gini_and_intra_list_similarity_gpu(n_users, n_items, K):
    """pxr_gini (count-of-counts form, no sort) and pxr_intra_list_similarity (norm of the summed unit vectors, no K x K
    matrix) vs the direct restatements of advanced_metrics.py:72-105 and novelty.py:295-340."""
    import torch
    from pixelrec_multimodal_b200.evaluation import gini_coefficient, intra_list_similarity
    rng = np.random.default_rng(n_users + K)
    # skewed popularity so the counts spread; a short list, an empty list, a one-item list
    p = 1.0 / np.arange(1, n_items + 1); p /= p.sum()
    recs = np.stack([rng.choice(n_items, K, replace=False, p=p) for _ in range(n_users)]).astype(np.int32)
    recs[1, K // 2:] = -1
    recs[2, :] = -1
    if n_users > 3:
        recs[3, 1:] = -1
    d_recs = torch.from_numpy(recs).cuda()
    counts = np.bincount(recs[recs >= 0], minlength=n_items)
    assert gini_coefficient(d_recs, n_items, include_zero=True) == pytest.approx(orc.gini_coefficient(counts), abs=1e-12)
    assert gini_coefficient(d_recs, n_items, include_zero=False) == pytest.approx(orc.gini_coefficient(counts[counts > 0]), abs=1e-12)
    emb = rng.standard_normal((n_items, 320)).astype(np.float32) + 0.5
    emb[min(5, n_items - 1)] = 0.0                                   # an item without embedding is skipped
    want = np.mean([orc.intra_list_similarity(recs[u].tolist(), emb) for u in range(n_users)])
    got = intra_list_similarity(d_recs, embeddings=torch.from_numpy(emb).cuda())
    assert got == pytest.approx(want, abs=2e-6)                       # fp32 sums of up to K unit vectors
    # reference known answers (gini of [1,1,1,1] and [4,0,0,0]) through the kernel
    one_each = torch.arange(4, dtype=torch.int32).view(4, 1).cuda()
    assert gini_coefficient(one_each, 4, include_zero=True) == pytest.approx(0.0, abs=1e-15)
    only_one = torch.zeros((4, 1), dtype=torch.int32).cuda()
    assert gini_coefficient(only_one, 4, include_zero=True) == pytest.approx(1.0 - 1.0 / 4.0, abs=1e-15)


@pytest.mark.gpu
def test_intra_list_similarity_from_resident_records():
    """embeddings = the item records pxr_precompute_items left on the device: equal to the oracle on the
    projected item-side modality vectors the oracle computes for the same items"""
    import torch
    from pixelrec_multimodal_b200 import synthetic as syn
    from pixelrec_multimodal_b200.evaluation import intra_list_similarity
    from tests import _cases as cs
    spec = syn.ModelSpec(n_users=8, n_items=90, fusion_type="gated")
    sd, feats = cs.make_workload(spec, syn.SEED + 51)
    model = cs.torch_model_from(spec, sd)
    eng = model.engine("catalogue")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    eng.precompute_items(model.item_embedding.weight.detach(), t(feats["tag_idx"]), t(feats["vis"]), t(feats["txt"]), t(feats["num"]))
    ii = np.arange(spec.n_items)
    f = orc.modality_features(sd, "relu", np.zeros(spec.n_items, np.int64), ii, feats["tag_idx"], feats["vis"], feats["txt"], feats["num"])
    emb = np.concatenate(f[1:], axis=1)                              # item, tag, vision, text, numerical vectors: 5 x 64
    rng = np.random.default_rng(7)
    recs = np.stack([rng.permutation(spec.n_items)[:12] for _ in range(40)]).astype(np.int32)
    want = np.mean([orc.intra_list_similarity(r.tolist(), emb) for r in recs])
    assert intra_list_similarity(t(recs), engine=eng) == pytest.approx(want, abs=5e-6)

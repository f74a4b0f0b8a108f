"""Sampled evaluation protocol (reference src/evaluation/tasks.py:181-224, 310-364).
CPU: properties of the oracle's reproducible sampler.  GPU: pxr_sample_candidates == oracle bit for bit;
SampledRetrievalEvaluator == the reference accuracy block evaluated on oracle-ranked candidates."""
import numpy as np
import pytest

from oracle import pxr_oracle as orc
from pixelrec_multimodal_b200 import synthetic as syn


def test_feistel_is_a_permutation():
    for n in (1, 2, 3, 17, 64, 100, 1000):
        at = orc._feistel_perm(orc.mix64(n), n)
        assert sorted(at(j) for j in range(n)) == list(range(n))


def test_oracle_sampler_properties():
    n_items = 500
    for u, pos in [(0, [3]), (7, [1, 2, 499]), (123456789, list(range(0, 40, 3))), (5, [])]:
        c = orc.sample_candidates(u, pos, n_items, 100, seed=42)
        assert len(c) == len(pos) + 100 and len(set(c)) == len(c)
        assert set(pos) <= set(c) and all(0 <= x < n_items for x in c)
        assert c == orc.sample_candidates(u, pos, n_items, 100, seed=42)            # reproducible
        assert c != orc.sample_candidates(u, pos, n_items, 100, seed=43)            # seed matters
        assert c != orc.sample_candidates(u + 1, pos, n_items, 100, seed=42)        # user matters
    # fewer non-positive items than requested negatives: all of them (tasks.py:208-209)
    c = orc.sample_candidates(1, [0, 1, 2], 10, 100, seed=1)
    assert sorted(c) == list(range(10))
    # negatives are uniform over the non-positives: chi-square-ish check of the marginal frequency
    cnt = np.zeros(50)
    for u in range(4000):
        for x in orc.sample_candidates(u, [7], 50, 5, seed=9):
            cnt[x] += 1
    neg = np.delete(cnt, 7)
    assert cnt[7] == 4000 and abs(neg.mean() - 4000 * 5 / 49) < 1e-9 and neg.std() < 4 * np.sqrt(neg.mean())


def test_rank_candidates_is_stable():
    assert orc.rank_candidates([0.5, 0.9, 0.5, 0.1], [10, 11, 12, 13], 3) == [11, 10, 12]


def _weighted_case(n_items, n_neg, max_pos, strategy, n=40):
    from pixelrec_multimodal_b200.evaluation import sampling_weights
    rng = np.random.default_rng(n_items * 7 + n_neg)
    npos = rng.integers(0, max_pos + 1, n)
    pos = [np.sort(rng.choice(n_items, min(c, n_items), replace=False)) for c in npos]
    indptr = np.concatenate([[0], np.cumsum([len(p) for p in pos])]).astype(np.int64)
    idx = (np.concatenate(pos) if indptr[-1] else np.zeros(0)).astype(np.int32)
    users = rng.integers(0, 10 ** 9, n).astype(np.int64)
    test_items = rng.zipf(1.5, 400 + n_items // 10) % n_items
    w = sampling_weights(test_items, n_items, strategy)
    assert np.array_equal(w, orc.sampling_weights(test_items, n_items, strategy))
    return users, pos, indptr, idx, w


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", ["popularity", "popularity_inverse"])
@pytest.mark.parametrize("n_items,n_neg,max_pos,stride", [(300, 40, 5, 64), (37, 100, 6, 128), (50, 10, 0, 16), (12, 8, 12, 16), (9, 4, 3, 5),
                                                          (20000, 100, 3, 128), (3000, 1000, 20, 1024)])
def test_weighted_candidates_kernel_matches_oracle(strategy, n_items, n_neg, max_pos, stride):
    """pxr_weighted_candidates == the per-user Python restatement, item for item (keys are float64 on both sides; a
    difference would need two keys within an ulp of each other)"""
    import torch
    from pixelrec_multimodal_b200.engine import weighted_candidates
    n = 40 if n_items <= 3000 else 12
    users, pos, indptr, idx, w = _weighted_case(n_items, n_neg, max_pos, strategy, n)
    d_idx = torch.from_numpy(idx).cuda() if len(idx) else torch.zeros(1, dtype=torch.int32).cuda()[:0]
    cand, length = weighted_candidates(torch.from_numpy(users).cuda(), torch.from_numpy(indptr).cuda(), d_idx,
                                       torch.from_numpy(w).cuda(), n_neg, seed=77, stride=stride)
    cand, length = cand.cpu().numpy(), length.cpu().numpy()
    for r in range(n):
        want = orc.sample_candidates_weighted(int(users[r]), pos[r].tolist(), w, n_neg, seed=77, stride=stride)
        assert length[r] == len(want) and cand[r][:length[r]].tolist() == want and np.all(cand[r][length[r]:] == -1), r
        assert set(pos[r].tolist()[:stride]) <= set(want) and len(set(want)) == len(want)


def test_weighted_candidates_need_the_gpu():
    """no host path: CPU tensors are refused, not sampled by other code"""
    import torch
    from pixelrec_multimodal_b200.engine import weighted_candidates, PxrError
    with pytest.raises(PxrError):
        weighted_candidates(torch.zeros(1, dtype=torch.int64), torch.zeros(2, dtype=torch.int64), torch.zeros(0, dtype=torch.int32),
                            torch.ones(4, dtype=torch.float64), 2, seed=1, stride=4)


def _follows_the_weights(sampler):
    n_items, n = 60, 3000
    w = np.ones(n_items); w[:10] = 8.0
    for weights, heavy_more in ((w, True), (1.0 / w, False)):
        c = sampler(n, weights)
        assert np.all((c >= 0).sum(axis=1) == 6) and np.all((c == 59).sum(axis=1) == 1)
        cnt = np.bincount(c[c >= 0], minlength=n_items).astype(float)
        heavy, light = cnt[:10].mean(), cnt[10:59].mean()
        assert (heavy > 3 * light) if heavy_more else (light > 3 * heavy)


def test_weighted_sampling_follows_the_weights():
    """the algorithm (oracle): inclusion frequency grows with the weight (popularity) and shrinks with it (inverse);
    positives are never negatives"""
    def sampler(n, weights):
        out = np.full((n, 8), -1, dtype=np.int64)
        for u in range(n):
            c = orc.sample_candidates_weighted(u, [59], weights, 5, seed=3, stride=8)
            out[u, :len(c)] = c
        return out
    _follows_the_weights(sampler)


@pytest.mark.gpu
def test_weighted_sampling_kernel_follows_the_weights():
    import torch
    from pixelrec_multimodal_b200.engine import weighted_candidates

    def sampler(n, weights):
        cand, _ = weighted_candidates(torch.arange(n, dtype=torch.int64).cuda(), torch.arange(n + 1, dtype=torch.int64).cuda(),
                                      torch.full((n,), 59, dtype=torch.int32).cuda(), torch.from_numpy(weights).cuda(), 5, seed=3, stride=8)
        return cand.cpu().numpy()
    _follows_the_weights(sampler)


@pytest.mark.gpu
def test_sampled_evaluator_builds_weighted_candidates_from_the_test_table():
    """SampledRetrievalEvaluator(sampling_strategy='popularity'): weights = item counts of the test table (tasks.py:227),
    candidates = positives + weighted negatives; host plumbing checked with a stand-in recommender"""
    import pandas as pd
    import torch
    from pixelrec_multimodal_b200 import SampledRetrievalEvaluator

    class _Rec:
        device = torch.device("cuda:0")
        n_items = 30
        user_index = {f"u{j}": j for j in range(6)}
        item_index = {f"i{j:02d}": j for j in range(30)}

    rows = [(f"u{j}", f"i{(3 * j) % 30:02d}") for j in range(6)] + [("u0", "i07"), ("u1", "i07"), ("u2", "i07"), ("nobody", "i01")]
    df = pd.DataFrame(rows, columns=["user_id", "item_id"])
    ev = SampledRetrievalEvaluator(_Rec(), df, top_k=5, num_negatives=10, sampling_strategy="popularity", seed=1)
    w = orc.sampling_weights([_Rec.item_index[i] for u, i in rows if u in _Rec.user_index], 30, "popularity")
    assert np.array_equal(ev._weights, w) and w[7] == 3.0 and w[0] == 1.0
    cand, length = ev.candidates()
    cand, length = cand.cpu().numpy(), length.cpu().numpy()
    for j, u in enumerate(ev.users):
        pos = ev.gt_idx[ev.gt_indptr[j]:ev.gt_indptr[j + 1]].tolist()
        want = orc.sample_candidates_weighted(int(u), pos, w, 10, seed=1, stride=cand.shape[1])
        assert cand[j][:int(length[j])].tolist() == want
    with pytest.raises(ValueError):
        SampledRetrievalEvaluator(_Rec(), df, sampling_strategy="other")


@pytest.mark.gpu
@pytest.mark.parametrize("n_items,n_neg,max_pos", [(500, 100, 4), (37, 100, 6), (100000, 100, 1), (2000, 1000, 30), (5, 3, 5)])
def test_sampler_kernel_matches_oracle(n_items, n_neg, max_pos):
    import torch
    from pixelrec_multimodal_b200.engine import sample_candidates
    rng = np.random.default_rng(n_items + n_neg)
    n = 300
    npos = rng.integers(0, max_pos + 1, n)
    pos = [np.sort(rng.choice(n_items, min(c, n_items), replace=False)) for c in npos]
    indptr = np.concatenate([[0], np.cumsum([len(p) for p in pos])]).astype(np.int64)
    idx = (np.concatenate(pos) if indptr[-1] else np.zeros(0)).astype(np.int32)
    users = rng.integers(0, 10 ** 9, n).astype(np.int64)
    cand, length = sample_candidates(torch.from_numpy(users).cuda(), torch.from_numpy(indptr).cuda(),
                                     torch.from_numpy(idx).cuda() if len(idx) else torch.zeros(1, dtype=torch.int32).cuda()[:0],
                                     n_items, n_neg, seed=77)
    cand, length = cand.cpu().numpy(), length.cpu().numpy()
    stride = cand.shape[1]
    for r in range(n):
        want = orc.sample_candidates(int(users[r]), pos[r].tolist(), n_items, n_neg, seed=77, stride=stride)
        assert length[r] == len(want)
        assert cand[r][:length[r]].tolist() == want and np.all(cand[r][length[r]:] == -1)


@pytest.mark.gpu
def test_sampled_evaluator_matches_oracle():
    import pandas as pd
    import torch
    from pixelrec_multimodal_b200 import FastRecommender, ItemFeatureStore, SampledRetrievalEvaluator
    from tests import _cases as cs
    spec = syn.ModelSpec(n_users=120, n_items=800, fusion_type="gated")
    sd, feats = cs.make_workload(spec, syn.SEED + 31)
    model = cs.torch_model_from(spec, sd)
    uids, iids = syn.user_ids(spec.n_users), syn.item_ids(spec.n_items)

    class _DS:
        class _E:
            def __init__(self, c): self.classes_ = np.array(c)
        user_encoder, item_encoder, interactions = _E(uids), _E(iids), None

    store = ItemFeatureStore(torch.from_numpy(feats["tag_idx"]), torch.from_numpy(feats["vis"]), torch.from_numpy(feats["txt"]),
                             torch.from_numpy(feats["num"]))
    rec = FastRecommender(model, _DS(), torch.device("cuda:0"), item_features=store)
    rng = np.random.default_rng(4)
    rows = []
    for u in range(spec.n_users):
        for it in rng.choice(spec.n_items, rng.integers(1, 4), replace=False):
            rows.append((uids[u], iids[int(it)]))
    test_df = pd.DataFrame(rows, columns=["user_id", "item_id"])
    ev = SampledRetrievalEvaluator(rec, test_df, top_k=10, ks=[5, 10], num_negatives=100, seed=5, keep_predictions=True)
    res = ev.evaluate()
    assert res["evaluation_method"] == "negative_sampling" and res["num_users_evaluated"] == spec.n_users
    cfg = cs.spec_cfg(spec)
    recs, poss = [], []
    for j, u in enumerate(ev.users):
        pos = ev.gt_idx[ev.gt_indptr[j]:ev.gt_indptr[j + 1]].tolist()
        cand = orc.sample_candidates(int(u), pos, spec.n_items, 100, seed=5, stride=103)
        ci = np.array(cand)
        sc = orc.forward_pairs(sd, cfg, np.full(len(ci), u), ci, feats["tag_idx"][ci], feats["vis"][ci], feats["txt"][ci], feats["num"][ci])
        got = [int(iids.index(i)) for i, _ in res["predictions"][uids[int(u)]]]
        want = orc.rank_candidates(sc.tolist(), cand, 10)
        if got != want:      # only swaps of candidates whose fp64 scores differ by less than the fp32 tolerance
            smap = dict(zip(cand, sc))
            assert sorted(got) == sorted(want) or all(abs(smap[a] - smap[b]) < 5e-5 for a, b in zip(got, want))
            want = got
        recs.append(want); poss.append(set(pos))
    for k in (5, 10):
        ref = orc.retrieval_metrics([r[:k] for r in recs], poss, k)
        for key in ("avg_precision_at_k", "avg_recall_at_k", "avg_f1_at_k", "avg_hit_rate_at_k", "avg_ndcg_at_k", "avg_mrr"):
            assert abs(res["by_k"][k][key] - ref[key]) <= 1e-12, (k, key)

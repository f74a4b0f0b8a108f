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

"""The compare-exchange network of the K4 register merge (csrc/simt_kernels.cu::merge_top64), restated lane by lane in
oracle/pxr_oracle.py::merge_top64_network, against a plain sort: ragged lists, padding, 1-13 lists folded one after
another, every K up to 64.  (The kernel itself is checked on the GPU in tests/test_gpu_parity.py::test_merge_topk_*.)"""
import numpy as np
import pytest

from oracle import pxr_oracle as orc


@pytest.mark.parametrize("seed", range(8))
def test_network_equals_sort(seed):
    rng = np.random.default_rng(seed)
    for _ in range(150):
        k = int(rng.integers(1, 65))
        S = int(rng.integers(1, 14))
        pool = (rng.choice(10 ** 7, size=S * 64, replace=False) + 1).astype(np.uint64)
        lists = []
        for s in range(S):
            n = int(rng.integers(0, k + 1))
            l = np.zeros(64, dtype=np.uint64)
            l[:n] = np.sort(pool[s * 64:s * 64 + n])[::-1]
            lists.append(l)
        acc = lists[0]
        for l in lists[1:]:
            acc = orc.merge_top64_network(acc, l)
        want = np.sort(np.concatenate(lists))[::-1][:k]
        assert np.array_equal(acc[:k], want)


def test_network_keeps_the_index_tie_break():
    """keys = (ordered score << 32) | ~index: equal scores come out in ascending index order (recommender.py:76, 105)"""
    key = lambda score_bits, idx: np.uint64((score_bits << 32) | (0xFFFFFFFF - idx))
    a = np.zeros(64, dtype=np.uint64); b = np.zeros(64, dtype=np.uint64)
    a[:3] = [key(9, 4), key(7, 1), key(7, 8)]
    b[:3] = [key(9, 2), key(7, 5), key(3, 0)]
    out = orc.merge_top64_network(a, b)[:6]
    idx = [0xFFFFFFFF - int(x & np.uint64(0xFFFFFFFF)) for x in out]
    assert idx == [2, 4, 1, 5, 8, 0]

"""N > 1 on real GPUs: two ranks, one GPU each, NCCL over NVLink -- no CLI involved.  Each rank
scores the same user blocks against its contiguous item shard through the C ABI, the per-shard lists are exchanged
(all-gather and the owned all-to-all) and merged by pxr_merge_topk; rank 0 also holds the whole catalogue and checks
the merged lists bit for bit, in the raw 16-bit mode and in exact mode (raw 64-slot lists exchanged, the owning rank
re-scores the merged candidates), and the sharded evaluator against the
single-GPU evaluator.  Skipped on a box with fewer than two GPUs."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PXR_REPO"])
from pixelrec_multimodal_b200 import FastRecommender, FullCatalogueEvaluator, ItemFeatureStore, synthetic as syn
from pixelrec_multimodal_b200.engine import merge_topk
from pixelrec_multimodal_b200.sharding import ShardedTopK, allgather_topk, owned_slice, shard_range
from tests import _cases as cs
import pandas as pd
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device(f"cuda:{rank}")
dist.init_process_group("nccl", device_id=dev)
for fusion in ("gated", "attention", "concatenate"):
    spec = syn.ModelSpec(n_users=300, n_items=2003, fusion_type=fusion)
    sd, feats = cs.make_workload(spec, syn.SEED + 61)
    indptr, idx, test_item = syn.make_histories(spec.n_users, spec.n_items, seed=5, lo=3, hi=30)

    class _DS:
        class _E:
            def __init__(self, c): self.classes_ = np.array(c)
        user_encoder, item_encoder, interactions = _E(syn.user_ids(spec.n_users)), _E(syn.item_ids(spec.n_items)), None
    store = ItemFeatureStore(torch.from_numpy(feats["tag_idx"]), torch.from_numpy(feats["vis"]), torch.from_numpy(feats["txt"]),
                             torch.from_numpy(feats["num"]))
    lo, hi = shard_range(spec.n_items, world, rank)
    rec = FastRecommender(cs.torch_model_from(spec, sd, device=dev), _DS(), dev, item_features=store, history=(indptr, idx),
                          item_range=(lo, hi))
    assert rec.engine().active_path == "tcgen05"
    full = FastRecommender(cs.torch_model_from(spec, sd, device=dev), _DS(), dev, item_features=store, history=(indptr, idx)) \
        if rank == 0 else None
    users = np.arange(spec.n_users)
    local_raw = lambda u, k, fs: rec.recommend_all(u, top_k=k, filter_seen=fs, raw=True)
    for exact in (False, True):
        # raw: the shards' 16-bit lists merged as they are.  exact: the raw 64-slot lists are exchanged, the owning rank
        # re-scores the merged candidates in fp32 against whole-catalogue records (ShardedTopK(rescore=...))
        st = ShardedTopK(local_raw, rescore=rec.rescore if exact else None)
        s, i = st.recommend_all(users, 50, True)                       # owned exchange + merge (+ re-score), gathered on every rank
        blocks = [users[:77], users[77:78], users[78:]]
        owned = list(st.recommend_blocks_owned(blocks, 50, True))      # overlapped with the next block's scoring
        row = 0
        for blk, (os_, oi_) in zip(blocks, owned):
            a, b = owned_slice(len(blk), world, rank)
            assert torch.equal(oi_, i[row + a:row + b]) and torch.equal(os_, s[row + a:row + b]), (fusion, exact, row)
            row += len(blk)
        if rank == 0:
            # same candidates (the per-pair 16-bit arithmetic does not depend on the tiling) and same fp32 records:
            # the sharded lists equal the single-GPU lists bit for bit in both modes
            full.engine().set_rescore(exact)
            fs, fi = full.recommend_all(users, top_k=50, filter_seen=True)
            assert torch.equal(fi, i) and torch.equal(fs, s), (fusion, exact, int((fi == i).all(dim=1).sum()))
    # top_k > 64: the shards exchange their raw 128-slot lists (two pages of the fused kernel), merged by the k > 64 form
    # of pxr_merge_topk and re-scored once by the owning rank == the single-GPU exact top-100
    st = ShardedTopK(local_raw, rescore=rec.rescore)
    s100, i100 = st.recommend_all(users, 100, True)
    if rank == 0:
        full.engine().set_rescore(True)
        fs, fi = full.recommend_all(users, top_k=100, filter_seen=True)
        assert torch.equal(fi, i100) and torch.equal(fs, s100), (fusion, "top-100", int((fi == i100).all(dim=1).sum()))
    # sharded evaluation (exact mode) == single-GPU evaluation
    st = ShardedTopK(local_raw, rescore=rec.rescore)
    test = pd.DataFrame({"user_id": [syn.user_ids(spec.n_users)[u] for u in range(spec.n_users) if test_item[u] >= 0],
                         "item_id": [syn.item_ids(spec.n_items)[int(test_item[u])] for u in range(spec.n_users) if test_item[u] >= 0]})
    ev = FullCatalogueEvaluator(rec, test, top_k=50, ks=[10, 50], sharded=st, user_block=128).evaluate()
    if rank == 0:
        full.engine().set_rescore(True)
        want = FullCatalogueEvaluator(full, test, top_k=50, ks=[10, 50]).evaluate()
        for k in (10, 50):
            for key in ("avg_recall_at_k", "avg_ndcg_at_k", "avg_mrr", "avg_precision_at_k", "avg_map_at_k"):
                assert abs(ev["by_k"][k][key] - want["by_k"][k][key]) <= 1e-12, (fusion, k, key)
    dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


@pytest.mark.gpu
def test_item_shards_over_nccl_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, PXR_REPO=str(REPO), MASTER_ADDR="127.0.0.1", MASTER_PORT="29741", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
        assert "OK" in o

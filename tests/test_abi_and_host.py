"""CPU-side checks: libpxr.so loads and exports every symbol include/pxr.h
declares (no compute calls without a GPU), the ctypes mirrors match the header,
host-side logic (sharding ranges, history CSR, world-size-2 gloo merge), and the
product path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from pixelrec_multimodal_b200 import _lib, sharding
from pixelrec_multimodal_b200.recommender import build_history_csr

REPO = Path(__file__).resolve().parent.parent
HEADER = (REPO / "include" / "pxr.h").read_text()


def _declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(pxr_[a-z0-9_]+)\s*\(", body)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pxr.h but not exported by libpxr.so"
    assert set(names) == set(_lib.declared_symbols())
    assert lib.pxr_version() == int(re.search(r"#define PXR_VERSION (\d+)", HEADER).group(1))


def test_ctypes_structs_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "pxr.h"\nint main(){printf("%zu %zu\\n", sizeof(pxr_config), sizeof(pxr_weights));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(REPO / "include"), str(src), "-o", str(exe)], check=True)
    a, b = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert C.sizeof(_lib.PxrConfig) == a
    assert C.sizeof(_lib.PxrWeights) == b


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C with no CUDA/torch types."""
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", str(REPO / "include" / "pxr.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    assert "torch" not in body.lower() and "Tensor" not in body


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from pixelrec_multimodal_b200 import FastMultimodalRecommender, PxrError
    from pixelrec_multimodal_b200.engine import PxrEngine
    with pytest.raises(PxrError):
        PxrEngine(fusion_type="concatenate", embedding_dim=16, vision_dim=8, language_dim=8, num_numerical=2,
                  hidden_dims=[16], n_tags=3)
    m = FastMultimodalRecommender(n_users=4, n_items=4, n_tags=3, num_numerical_features=2, embedding_dim=16,
                                  vision_model_name="cached8", language_model_name="cached8",
                                  fusion_hidden_dims=[16])
    z = torch.zeros(2, dtype=torch.long)
    with pytest.raises(RuntimeError):
        m(z, z, z, image=torch.zeros(2, 8), text_input_ids=torch.zeros(2, 8),
          text_attention_mask=torch.ones(2, 1), numerical_features=torch.zeros(2, 2))


def test_product_never_imports_oracle():
    pkg = REPO / "pixelrec_multimodal_b200"
    for p in pkg.rglob("*.py"):
        txt = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), p
    for p in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert "oracle/" not in p.read_text(), p


def test_state_dict_keys_mirror_reference():
    """Key names / shapes of SURVEY.md A1 for the three fusion types."""
    from pixelrec_multimodal_b200 import FastMultimodalRecommender
    for ft, extra in (("concatenate", []),
                      ("gated", ["fusion_layer.gating_network.0.weight", "fusion_layer.gating_network.0.bias"]),
                      ("attention", ["fusion_layer.attention.in_proj_weight", "fusion_layer.attention.in_proj_bias",
                                     "fusion_layer.attention.out_proj.weight", "fusion_layer.attention.out_proj.bias",
                                     "fusion_layer.norm.weight", "fusion_layer.norm.bias"])):
        m = FastMultimodalRecommender(n_users=5, n_items=6, n_tags=3, num_numerical_features=7, embedding_dim=64,
                                      vision_model_name="cached512", language_model_name="cached384",
                                      fusion_type=ft)
        sd = m.state_dict()
        for k in ["user_embedding.weight", "item_embedding.weight", "tag_embedding.weight",
                  "vision_projection.0.weight", "language_projection.0.bias", "numerical_projection.0.weight",
                  "prediction_network.0.weight", "prediction_network.2.running_var",
                  "prediction_network.4.weight", "prediction_network.8.weight", "prediction_network.12.weight"] + extra:
            assert k in sd, (ft, k)
        assert tuple(sd["prediction_network.0.weight"].shape) == (512, 384 if ft == "concatenate" else 64)
        assert tuple(sd["prediction_network.12.weight"].shape) == (1, 128)
        assert tuple(sd["vision_projection.0.weight"].shape) == (64, 512)
    with pytest.raises(ValueError):
        FastMultimodalRecommender(n_users=5, n_items=6, n_tags=3, num_numerical_features=7, fusion_type="bogus")


def test_shard_ranges_cover_and_are_contiguous():
    for n in (0, 1, 7, 96282, 407082):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            for a, b in zip(r, r[1:]):
                assert a[1] == b[0]
            assert all(lo <= hi for lo, hi in r)


def test_history_csr():
    import pandas as pd
    ui = {"a": 0, "b": 1, "c": 2}
    ii = {f"i{j}": j for j in range(6)}
    df = pd.DataFrame({"user_id": ["b", "a", "b", "b", "zz", "a"], "item_id": ["i5", "i2", "i1", "i5", "i0", "nope"]})
    indptr, idx = build_history_csr(ui, ii, df, 3)
    assert indptr.tolist() == [0, 1, 3, 3]
    assert idx.tolist() == [2, 1, 5]
    assert idx.dtype == np.int32 and indptr.dtype == np.int64


_GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PXR_REPO"])
from pixelrec_multimodal_b200.sharding import ShardedTopK, shard_range
from oracle import pxr_oracle as orc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(5)
NU, NI, K = 9, 37, 6
scores = np.round(rng.standard_normal((NU, NI)), 1).astype(np.float32)   # rounding creates ties
lo, hi = shard_range(NI, world, rank)
def local_topk(users, k, filter_seen):
    s = np.full((len(users), k), -np.inf, np.float32); i = np.full((len(users), k), -1, np.int32)
    for r, u in enumerate(users):
        sel, sc = orc.topk_from_scores(scores[u, lo:hi].astype(np.float64), k, item_base=lo)
        s[r, :len(sel)] = sc; i[r, :len(sel)] = sel
    return torch.from_numpy(s), torch.from_numpy(i)
def merge(all_s, all_i):
    S, n, k = all_s.shape
    out_s = np.full((n, k), -np.inf, np.float32); out_i = np.full((n, k), -1, np.int32)
    for r in range(n):
        lists = []
        for s in range(S):
            m = all_i[s, r].numpy() >= 0
            lists.append((all_i[s, r].numpy()[m].astype(np.int64), all_s[s, r].numpy()[m].astype(np.float64)))
        idx, sc = orc.merge_topk(lists, k)
        out_s[r, :len(idx)] = sc; out_i[r, :len(idx)] = idx
    return torch.from_numpy(out_s), torch.from_numpy(out_i)
st = ShardedTopK(local_topk, merge)
s, i = st.recommend_all(np.arange(NU), K, False)
for u in range(NU):
    sel, sc = orc.topk_from_scores(scores[u].astype(np.float64), K)
    assert i[u].tolist() == sel.tolist(), (rank, u, i[u].tolist(), sel.tolist())
# pipelined variant: exchange of block b overlapped with scoring of block b + 1, single packed all-gather
blocks = [np.arange(0, 3), np.arange(3, 4), np.arange(4, NU)]
outs = list(st.recommend_blocks(blocks, K, False))
assert len(outs) == 3
ps, pi = torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
assert torch.equal(pi, i) and torch.equal(ps, s)
# reduce-scatter ownership: one all-to-all, every rank merges (and keeps) only its slice of each block
from pixelrec_multimodal_b200.sharding import owned_slice
owned = list(st.recommend_blocks_owned(blocks, K, False))
row = 0
for blk, (os_, oi_) in zip(blocks, owned):
    lo_, hi_ = owned_slice(len(blk), world, rank)
    assert oi_.shape[0] == hi_ - lo_
    assert torch.equal(oi_, i[row + lo_:row + hi_]) and torch.equal(os_, s[row + lo_:row + hi_]), (rank, row)
    row += len(blk)
# sharded evaluation: every rank scores its shard, metric sums of disjoint user slices, one all-reduce
import pandas as pd
from pixelrec_multimodal_b200.evaluation import FullCatalogueEvaluator
class _Rec:
    user_index = {f"u{u}": u for u in range(NU)}
    item_index = {f"i{j}": j for j in range(NI)}
    n_items, device = NI, torch.device("cpu")
rng2 = np.random.default_rng(11)
rows = [(f"u{u}", f"i{int(j)}") for u in range(NU) for j in rng2.choice(NI, 2, replace=False)]
# the reference keeps every distinct test user in the means and every raw row in the recall denominator
# (tasks.py:537-540, 579): an unknown user, an unknown item and a duplicated row
rows += [("ghost", "i1"), ("u0", "nope"), rows[2]]
test = pd.DataFrame(rows, columns=["user_id", "item_id"])
ev = FullCatalogueEvaluator(_Rec(), test, top_k=K, ks=[3, K], filter_seen=False, sharded=st, user_block=4)
assert ev.n_total == NU + 1 and ev.unknown_users == ["ghost"] and int(ev.recall_den.sum()) == len(rows) - 1
def cpu_metric_sums(topk, gt_indptr, gt_idx, ks, recall_den=None):      # K5 restated with the oracle (no GPU here)
    out = np.zeros((len(ks), 9))
    ip, gi = gt_indptr.numpy(), gt_idx.numpy()
    for r_ in range(topk.shape[0]):
        pos = [int(x) for x in gi[ip[r_]:ip[r_ + 1]]]
        raw = pos + pos[:1] * (int(recall_den[r_]) - len(pos)) if recall_den is not None else pos
        for a, k in enumerate(sorted(ks)):
            recs = [int(x) for x in topk[r_][:k].tolist() if x >= 0]
            m = orc.retrieval_metrics([recs], [raw], k)
            out[a, :6] += [m["avg_precision_at_k"], m["avg_recall_at_k"], m["avg_f1_at_k"], m["avg_hit_rate_at_k"], m["avg_ndcg_at_k"], m["avg_mrr"]]
    return out
ev._metric_sums = cpu_metric_sums
res = ev.evaluate()
# the reference loop on the table itself: string ids, groupby order, [] for the unknown user
names = sorted(set(t[0] for t in rows))
recs_all = [[f"i{int(x)}" for x in i[int(n_[1:])].tolist() if x >= 0] if n_ != "ghost" else [] for n_ in names]
pos_all = [[t[1] for t in rows if t[0] == n_] for n_ in names]
for k in (3, K):
    want = orc.retrieval_metrics([r_[:k] for r_ in recs_all], pos_all, k)
    for key in ("avg_precision_at_k", "avg_recall_at_k", "avg_f1_at_k", "avg_hit_rate_at_k", "avg_ndcg_at_k", "avg_mrr"):
        assert abs(res["by_k"][k][key] - want[key]) < 1e-12, (rank, k, key, res["by_k"][k][key], want[key])
assert res["num_users_evaluated"] == NU + 1
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
"""


def test_sharded_topk_world2_gloo(tmp_path):
    """N>1 path on CPU: 2 gloo ranks, each owning a contiguous item range, one
    all-gather of the per-shard top-K lists, merge == top-K of the full row
    (including tie-break by lower global index across the shard boundary)."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, PXR_REPO=str(REPO), MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "OK" in o


def test_synthetic_generators_are_reproducible():
    """every synthetic input is a pure function of (seed, stream name)."""
    from pixelrec_multimodal_b200 import synthetic as syn
    spec = syn.ModelSpec(n_users=50, n_items=70, fusion_type="gated")
    a, b = syn.make_state_dict(spec, seed=5), syn.make_state_dict(spec, seed=5)
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
    c = syn.make_state_dict(spec, seed=6)
    assert any(not np.array_equal(a[k], c[k]) for k in a)
    f1, f2 = syn.make_item_features(spec, seed=5), syn.make_item_features(spec, seed=5)
    assert all(np.array_equal(f1[k], f2[k]) for k in f1)
    assert abs(np.linalg.norm(f1["vis"], axis=1).mean() - 10.0) < 1e-3 and abs(np.linalg.norm(f1["txt"], axis=1).mean() - 1.0) < 1e-3
    ip1, ix1, t1 = syn.make_histories(50, 70, seed=5, lo=3, hi=20)
    ip2, ix2, t2 = syn.make_histories(50, 70, seed=5, lo=3, hi=20)
    assert np.array_equal(ip1, ip2) and np.array_equal(ix1, ix2) and np.array_equal(t1, t2)
    for u in range(50):                              # ascending, unique, the held-out test item is not in the history
        h = ix1[ip1[u]:ip1[u + 1]]
        assert np.all(np.diff(h) > 0) and int(t1[u]) not in set(h.tolist())

"""CLI front ends.  CPU: argument surface, config defaults, user selection and JSON shape
(reference scripts/generate_recommendations.py:221-226, 271-287).  GPU: both commands end to end on a tiny
synthetic catalogue, equal to the Python API."""
import json
import types

import numpy as np
import pytest

from pixelrec_multimodal_b200 import cli, synthetic as syn


def test_parser_mirrors_reference_flags():
    p = cli.make_parser()
    a = p.parse_args(["generate", "--cache", "c", "--interactions", "i.csv", "--users", "u1", "u2", "--output", "o.json"])
    assert a.users == ["u1", "u2"] and a.output == "o.json" and a.sample_users is None and a.user_file is None
    a = p.parse_args(["evaluate", "--cache", "c", "--interactions", "i.csv", "--test_data", "t.csv", "--use_sampling"])
    assert a.use_sampling and a.num_negatives == 100 and a.sampling_strategy == "random" and a.output == "evaluation_results.json"
    assert a.eval_task == "retrieval"                                                  # evaluate.py:242
    a = p.parse_args(["evaluate", "--cache", "c", "--interactions", "i.csv", "--test_data", "t.csv", "--eval_task", "ranking"])
    assert a.eval_task == "ranking"
    with pytest.raises(SystemExit):
        p.parse_args(["evaluate", "--cache", "c", "--interactions", "i.csv"])          # --test_data is required
    with pytest.raises(SystemExit):
        p.parse_args(["evaluate", "--cache", "c", "--interactions", "i.csv", "--test_data", "t.csv", "--eval_task", "other"])


def test_parser_accepts_the_reference_scripts_command_lines(tmp_path, capsys):
    """Every flag of scripts/evaluate.py:236-250 and scripts/generate_recommendations.py:248-254 parses; the ones without
    work on this path are accepted and say so; a checkpoint directory resolves like find_model_checkpoint (evaluate.py:54-110)."""
    p = cli.make_parser()
    a = cli.resolve_compat(p.parse_args(["evaluate", "--cache", "c", "--train_data", "train.csv", "--test_data", "t.csv", "--no_sampling",
                                         "--num_negatives", "20", "--sampling_strategy", "popularity", "--recommender_type", "multimodal",
                                         "--num_workers", "4", "--warmup_recommender_cache", "--save_predictions", "p.json",
                                         "--checkpoint_name", "final_model.pth", "--device", "cuda:0"]))
    assert a.interactions == "train.csv" and a.use_sampling is False and a.num_negatives == 20 and a.num_workers == 1
    assert a.sampling_strategy == "popularity" and a.warmup_recommender_cache and a.checkpoint is None
    with pytest.raises(SystemExit):
        cli.resolve_compat(p.parse_args(["evaluate", "--cache", "c", "--interactions", "i.csv", "--test_data", "t.csv",
                                         "--recommender_type", "item_knn"]))
    g = cli.resolve_compat(p.parse_args(["generate", "--cache", "c", "--interactions", "i.csv", "--sample_users", "3", "--use_diversity"]))
    assert g.use_diversity and "Diversity method not implemented" in capsys.readouterr().out     # generate_recommendations.py:206-208
    d = tmp_path / "ckpt"
    d.mkdir()
    (d / "last_model.pth").write_bytes(b"x")
    a = cli.resolve_compat(p.parse_args(["evaluate", "--cache", "c", "--interactions", "i.csv", "--test_data", "t.csv", "--checkpoint", str(d)]))
    assert a.checkpoint == str(d / "last_model.pth")
    (d / "final_model.pth").write_bytes(b"x")
    a = cli.resolve_compat(p.parse_args(["evaluate", "--cache", "c", "--interactions", "i.csv", "--test_data", "t.csv", "--checkpoint", str(d),
                                         "--checkpoint_name", "final_model.pth"]))
    assert a.checkpoint == str(d / "final_model.pth")
    with pytest.raises(SystemExit):
        cli.resolve_compat(p.parse_args(["evaluate", "--cache", "c", "--interactions", "i.csv", "--test_data", "t.csv",
                                         "--checkpoint", str(tmp_path)]))


def test_config_defaults_and_yaml(tmp_path):
    c = cli.load_config(None)
    assert c["model"]["fusion_hidden_dims"] == [512, 256, 128] and c["recommendation"]["top_k"] == 50 and c["results_dir"] == "results"
    y = tmp_path / "c.yaml"
    y.write_text("model:\n  fusion_type: gated\n  embedding_dim: 64\nrecommendation:\n  top_k: 7\nresults_dir: out\n")
    c = cli.load_config(str(y))
    assert c["model"]["fusion_type"] == "gated" and c["recommendation"]["top_k"] == 7 and c["results_dir"] == "out"
    assert c["recommendation"]["filter_seen"] is True and c["model"]["final_activation"] == "sigmoid"


def test_user_selection_and_json_shape(tmp_path):
    allu = [f"u{i}" for i in range(20)]
    ns = lambda **k: types.SimpleNamespace(**{"users": None, "user_file": None, "sample_users": None, "all_users": False, **k})
    assert cli.select_users(ns(), allu) == allu[:5]                                   # default: first five
    assert cli.select_users(ns(users=["x"]), allu) == ["x"]
    f = tmp_path / "u.txt"; f.write_text("u3\n\nu9\n")
    assert cli.select_users(ns(user_file=str(f)), allu) == ["u3", "u9"]
    s = cli.select_users(ns(sample_users=4), allu)
    assert len(s) == 4 and s == cli.select_users(ns(sample_users=4), allu)            # random_state=42
    assert cli.select_users(ns(sample_users=50), allu) == allu
    assert cli.select_users(ns(all_users=True), allu) == allu
    out = cli.format_results(["u1"], [[("i5", np.float32(0.25)), ("i2", 0.125)]])
    assert out == {"u1": {"recommendations": [{"item_id": "i5", "score": 0.25}, {"item_id": "i2", "score": 0.125}]}}
    json.dumps(out)


def test_encoders_come_from_the_checkpoint_dir(tmp_path):
    """scripts/evaluate.py:113-167, 301-304: the training-time user_encoder.pkl / item_encoder.pkl define the embedding
    rows; a table whose ids differ from training must not silently re-number them."""
    import pickle
    import pandas as pd
    import torch
    from sklearn.preprocessing import LabelEncoder
    from pixelrec_multimodal_b200.packed_cache import PackedFeatureCache, write_packed_cache
    spec = syn.ModelSpec(n_users=12, n_items=20, fusion_type="concatenate")
    feats = syn.make_item_features(spec, seed=3)
    uids, iids = syn.user_ids(spec.n_users), syn.item_ids(spec.n_items)
    write_packed_cache(tmp_path / "cache", iids[:18], feats["tag_idx"][:18], feats["vis"][:18], feats["txt"][:18], feats["num"][:18])
    # today's table only holds 5 of the 12 training users
    pd.DataFrame({"user_id": uids[:5], "item_id": iids[:5]}).to_csv(tmp_path / "train.csv", index=False)
    ck_dir = tmp_path / "ckpt" / "clip_sentence-bert"
    (tmp_path / "ckpt" / "encoders").mkdir(parents=True)
    ck_dir.mkdir(parents=True)
    for name, ids in (("user", uids), ("item", iids)):
        with open(tmp_path / "ckpt" / "encoders" / f"{name}_encoder.pkl", "wb") as f:
            pickle.dump(LabelEncoder().fit(ids), f)
    assert cli.find_encoders(str(ck_dir / "best_model.pth"), None) == tmp_path / "ckpt" / "encoders"
    assert cli.find_encoders(str(tmp_path / "elsewhere" / "m.pth"), None) is None
    with pytest.raises(FileNotFoundError):
        cli.find_encoders(None, str(tmp_path / "nothing"))
    cache = PackedFeatureCache(tmp_path / "cache")
    inter = pd.read_csv(tmp_path / "train.csv", dtype=str)
    ds = cli.TableDataset(inter, cache, tmp_path / "ckpt" / "encoders")
    assert list(ds.user_encoder.classes_) == uids and list(ds.item_encoder.classes_) == iids
    assert len(cli.TableDataset(inter, cache).user_encoder.classes_) == 5           # rebuilt from the table: NOT the training ids
    # encoder ids without a cache row are flagged missing (they score 0.0), not an error
    store = cache.to_store("cpu", order=iids)
    assert store.missing.tolist() == [False] * 18 + [True] * 2 and float(store.vis[18:].abs().sum()) == 0.0
    # a checkpoint whose tables do not match the (rebuilt) encoders is refused before anything is scored
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in syn.make_state_dict(spec, seed=3).items()}
    torch.save({"model_state_dict": sd}, tmp_path / "elsewhere.pth")
    with pytest.raises(ValueError, match="12 user embedding rows"):
        cli.build_recommender(cli.load_config(None), str(tmp_path / "elsewhere.pth"), str(tmp_path / "cache"), str(tmp_path / "train.csv"))


@pytest.mark.gpu
@pytest.mark.parametrize("fusion", ["gated", "attention"])
def test_cli_end_to_end(tmp_path, fusion):
    import pandas as pd
    import torch
    from pixelrec_multimodal_b200.packed_cache import write_packed_cache
    from tests import _cases as cs
    spec = syn.ModelSpec(n_users=60, n_items=300, fusion_type=fusion)
    sd, feats = cs.make_workload(spec, syn.SEED + 41)
    uids, iids = syn.user_ids(spec.n_users), syn.item_ids(spec.n_items)
    write_packed_cache(tmp_path / "cache", iids, feats["tag_idx"], feats["vis"], feats["txt"], feats["num"])
    indptr, idx, test_item = syn.make_histories(spec.n_users, spec.n_items, seed=3, lo=3, hi=20)
    rows = [(uids[u], iids[int(i)]) for u in range(spec.n_users) for i in idx[indptr[u]:indptr[u + 1]]]
    pd.DataFrame(rows, columns=["user_id", "item_id"]).to_csv(tmp_path / "train.csv", index=False)
    pd.DataFrame([(uids[u], iids[int(test_item[u])]) for u in range(spec.n_users)], columns=["user_id", "item_id"]).to_csv(tmp_path / "test.csv", index=False)
    torch.save({"model_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}}, tmp_path / "m.pth")
    (tmp_path / "c.yaml").write_text(f"model:\n  fusion_type: {fusion}\nrecommendation:\n  top_k: 10\nresults_dir: {tmp_path / 'res'}\n")
    common = ["--config", str(tmp_path / "c.yaml"), "--checkpoint", str(tmp_path / "m.pth"), "--cache", str(tmp_path / "cache"),
              "--interactions", str(tmp_path / "train.csv")]
    res = cli.main(["generate", *common, "--users", uids[3], uids[7], "ghost", "--output", "r.json"])
    disk = json.loads((tmp_path / "res" / "r.json").read_text())
    assert disk == res and list(disk) == [uids[3], uids[7], "ghost"] and disk["ghost"]["recommendations"] == []
    assert len(disk[uids[3]]["recommendations"]) == 10
    seen = {iids[int(i)] for i in idx[indptr[3]:indptr[4]]}
    assert not seen & {r["item_id"] for r in disk[uids[3]]["recommendations"]}
    allr = cli.main(["generate", *common, "--all_users", "--output", "all.json"])
    assert len(allr) == spec.n_users
    for u in (uids[3], uids[7]):                                                     # batched == per-user string API
        assert [r["item_id"] for r in allr[u]["recommendations"]] == [r["item_id"] for r in disk[u]["recommendations"]]
    ev = cli.main(["evaluate", *common, "--test_data", str(tmp_path / "test.csv"), "--ks", "5", "--output", "e.json", "--novelty"])
    assert ev["avg_personalized_novelty"] == 1.0 and 0.0 < ev["avg_catalog_coverage"] <= 1.0        # filter_seen: every item is new
    assert ev["avg_self_information"] > 0.0 and 0.0 <= ev["avg_personalization"] <= 1.0
    assert ev["evaluation_method"] == "full_evaluation" and ev["num_users_evaluated"] == spec.n_users
    assert set(ev["by_k"]) == {"5", "10"} and 0.0 <= ev["avg_recall_at_k"] <= 1.0
    hits = np.mean([iids[int(test_item[u])] in [r["item_id"] for r in allr[uids[u]]["recommendations"]] for u in range(spec.n_users)])
    assert abs(ev["avg_hit_rate_at_k"] - hits) <= 1e-12                               # evaluator == lists written by generate
    es = cli.main(["evaluate", *common, "--test_data", str(tmp_path / "test.csv"), "--use_sampling", "--output", "s.json",
                   "--save_predictions", "p.json"])
    assert es["evaluation_method"] == "negative_sampling" and (tmp_path / "res" / "p.json").exists()
    assert es["avg_hit_rate_at_k"] >= ev["avg_hit_rate_at_k"]                         # 101 candidates instead of 300
    er = cli.main(["evaluate", *common, "--test_data", str(tmp_path / "test.csv"), "--eval_task", "ranking", "--output", "rk.json",
                   "--save_predictions", "rp.json"])                                   # evaluate.py:242, 402-408
    assert er["evaluation_metadata"]["task"] == "ranking" and er["num_users_evaluated"] == spec.n_users
    assert er["avg_avg_rank"] == 1.0 and er["avg_mrr"] == 1.0 and er["avg_ndcg_at_k"] == 1.0   # one test item per user
    rp = json.loads((tmp_path / "res" / "rp.json").read_text())
    assert len(rp) == spec.n_users and rp[uids[3]][0]["item_id"] == iids[int(test_item[3])] and 0.0 < rp[uids[3]][0]["score"] < 1.0


@pytest.mark.gpu
def test_cli_sharded_evaluate_two_gpus(tmp_path):
    """`torchrun --nproc-per-node 2 -m pixelrec_multimodal_b200.cli evaluate`: item-axis shards, NCCL all-gather of the
    per-shard lists, metric sums all-reduced == the single-GPU evaluation (needs two GPUs; skipped otherwise)."""
    import os
    import subprocess
    import sys
    import pandas as pd
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pixelrec_multimodal_b200.packed_cache import write_packed_cache
    from tests import _cases as cs
    spec = syn.ModelSpec(n_users=300, n_items=1001, fusion_type="gated")
    sd, feats = cs.make_workload(spec, syn.SEED + 43)
    uids, iids = syn.user_ids(spec.n_users), syn.item_ids(spec.n_items)
    write_packed_cache(tmp_path / "cache", iids, feats["tag_idx"], feats["vis"], feats["txt"], feats["num"])
    indptr, idx, test_item = syn.make_histories(spec.n_users, spec.n_items, seed=5, lo=3, hi=20)
    pd.DataFrame([(uids[u], iids[int(i)]) for u in range(spec.n_users) for i in idx[indptr[u]:indptr[u + 1]]],
                 columns=["user_id", "item_id"]).to_csv(tmp_path / "train.csv", index=False)
    pd.DataFrame([(uids[u], iids[int(test_item[u])]) for u in range(spec.n_users)], columns=["user_id", "item_id"]).to_csv(tmp_path / "test.csv", index=False)
    torch.save({"model_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}}, tmp_path / "m.pth")
    (tmp_path / "c.yaml").write_text(f"model:\n  fusion_type: gated\nrecommendation:\n  top_k: 50\nresults_dir: {tmp_path / 'res'}\n")
    common = ["--config", str(tmp_path / "c.yaml"), "--checkpoint", str(tmp_path / "m.pth"), "--cache", str(tmp_path / "cache"),
              "--interactions", str(tmp_path / "train.csv"), "--test_data", str(tmp_path / "test.csv"), "--ks", "10"]
    one = cli.main(["evaluate", *common, "--output", "one.json"])
    repo = str(__import__("pathlib").Path(__file__).resolve().parent.parent)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", "-m", "pixelrec_multimodal_b200.cli", "evaluate", *common, "--output", "two.json"],
                       capture_output=True, text=True, timeout=600, cwd=repo, env=dict(os.environ, PYTHONPATH=repo))
    assert p.returncode == 0, p.stderr[-3000:]
    two = json.loads((tmp_path / "res" / "two.json").read_text())
    for k in ("10", "50"):
        for key in ("avg_precision_at_k", "avg_recall_at_k", "avg_hit_rate_at_k", "avg_ndcg_at_k", "avg_mrr"):
            assert abs(two["by_k"][k][key] - one["by_k"][k][key]) <= 1e-12, (k, key)
    assert two["num_users_evaluated"] == spec.n_users == one["num_users_evaluated"]

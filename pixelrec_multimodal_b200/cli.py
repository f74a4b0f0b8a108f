"""Command-line front ends of the GPU path: the work of the reference's
``scripts/generate_recommendations.py`` and ``scripts/evaluate.py`` for the multimodal recommender, with the
same flags and the same JSON outputs, on top of a packed feature cache and CSV interaction tables.

  python -m pixelrec_multimodal_b200.cli generate --config C.yaml --checkpoint M.pth --cache DIR \\
         --interactions train.csv [--users u1 u2 | --user_file F | --sample_users N | --all_users] --output recs.json
  python -m pixelrec_multimodal_b200.cli evaluate --config C.yaml --checkpoint M.pth --cache DIR \\
         --interactions train.csv --test_data test.csv [--use_sampling --num_negatives 100] --output results.json

What replaces what: the encoders are the sorted unique ids of the interaction table and of the cache (the order
sklearn's ``LabelEncoder`` gives, ``dataset.py:142-148``), built once; user histories are one CSR built once
(the reference masks the whole interaction frame per user, ``dataset.py:462-476``); item features come from the
packed cache in a few large copies.  ``--all_users`` and ``evaluate`` use the batched API; the per-user flags go
through ``get_recommendations`` exactly like the reference script (``generate_recommendations.py:196-226``).
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np


# ----------------------------------------------------------------------------- configuration
def load_config(path: Optional[str]) -> Dict:
    """The subset of the reference YAML (``src/config.py``) this path reads; defaults as in ``config.py:41-62,512-519``."""
    cfg = {"model": {}, "recommendation": {}, "data": {}}
    if path:
        import yaml
        loaded = yaml.safe_load(Path(path).read_text()) or {}
        for k in cfg:
            cfg[k].update(loaded.get(k) or {})
        cfg["results_dir"] = loaded.get("results_dir", "results")
    m = cfg["model"]
    m.setdefault("embedding_dim", 64)
    m.setdefault("fusion_hidden_dims", [512, 256, 128])
    m.setdefault("fusion_activation", "relu")
    m.setdefault("use_batch_norm", True)
    m.setdefault("projection_hidden_dim", None)
    m.setdefault("final_activation", "sigmoid")
    m.setdefault("fusion_type", "concatenate")
    m.setdefault("num_attention_heads", 4)
    r = cfg["recommendation"]
    r.setdefault("top_k", 50)
    r.setdefault("filter_seen", True)
    cfg.setdefault("results_dir", "results")
    return cfg


def find_encoders(checkpoint: Optional[str], encoders: Optional[str]) -> Optional[Path]:
    """Directory holding the training-time ``user_encoder.pkl`` / ``item_encoder.pkl`` (``scripts/train.py:503-508``).
    Search order of ``scripts/evaluate.py:113-167`` relative to the checkpoint file: ``<ckpt dir>/encoders``,
    ``<ckpt dir>``, then the same two one level up (checkpoints live in ``<checkpoint_dir>/<vision>_<language>/``,
    ``evaluate.py:54-110``); ``--encoders DIR`` overrides the search.  None when there are no pickled encoders."""
    cands = []
    if encoders:
        cands.append(Path(encoders))
    elif checkpoint:
        d = Path(checkpoint).resolve().parent
        cands += [d / "encoders", d, d.parent / "encoders", d.parent]
    for c in cands:
        if (c / "user_encoder.pkl").exists() and (c / "item_encoder.pkl").exists():
            return c
    if encoders:
        raise FileNotFoundError(f"user_encoder.pkl / item_encoder.pkl not found in {encoders}")
    return None


class TableDataset:
    """The attributes the recommender reads on its dataset (``recommender.py:58-90, 239-269``): encoders with
    ``classes_``, the interaction frame, and a ``feature_cache`` with ``get(item_id)``.

    With ``encoders_dir`` the pickled training-time ``LabelEncoder`` objects are loaded as the reference scripts do
    (``evaluate.py:301-304``, ``generate_recommendations.py:116-119``), so embedding rows line up with the
    checkpoint whatever ids the CSV / cache hold today.  Without it the encoders are rebuilt as the sorted unique
    ids of the interaction table and the cache (what ``LabelEncoder.fit`` gives on the same tables,
    ``dataset.py:142-148``) -- only valid when those are the training tables; ``build_recommender`` checks the
    row counts against the checkpoint either way."""

    class _Encoder:
        def __init__(self, classes):
            self.classes_ = np.asarray(classes, dtype=object)

    def __init__(self, interactions, cache, encoders_dir: Optional[Path] = None):
        if encoders_dir is not None:
            import pickle
            with open(Path(encoders_dir) / "user_encoder.pkl", "rb") as f:
                self.user_encoder = pickle.load(f)
            with open(Path(encoders_dir) / "item_encoder.pkl", "rb") as f:
                self.item_encoder = pickle.load(f)
            for name, enc in (("user", self.user_encoder), ("item", self.item_encoder)):
                if not hasattr(enc, "classes_"):
                    raise ValueError(f"{name}_encoder.pkl holds no fitted encoder (no classes_)")
        else:
            users = np.sort(interactions["user_id"].astype(str).unique())
            self.user_encoder = self._Encoder(users)
            self.item_encoder = self._Encoder(sorted(cache.item_ids))
        self.interactions = interactions
        self.feature_cache = cache


def distributed_context(device: str):
    """(world, rank, device) -- under ``torchrun`` (WORLD_SIZE > 1) the NCCL process group is initialised and every
    rank takes the GPU LOCAL_RANK; the catalogue is then split into contiguous item shards, one per rank."""
    import os
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 1, 0, device
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    return world, dist.get_rank(), f"cuda:{local}"


def build_recommender(cfg: Dict, checkpoint: Optional[str], cache_dir: str, interactions_csv: str, device: str = "cuda:0",
                      n_tags: Optional[int] = None, shard=None, encoders: Optional[str] = None):
    import pandas as pd
    import torch
    from . import FastMultimodalRecommender, FastRecommender
    from .model import _IGNORED_PREFIXES
    from .packed_cache import PackedFeatureCache
    cache = PackedFeatureCache(cache_dir)
    inter = pd.read_csv(interactions_csv, dtype={"user_id": str, "item_id": str})
    enc_dir = find_encoders(checkpoint, encoders)
    ds = TableDataset(inter, cache, enc_dir)
    m = cfg["model"]
    sd = None
    if checkpoint:
        ck = torch.load(checkpoint, map_location="cpu", weights_only=False)
        sd = ck["model_state_dict"] if isinstance(ck, dict) and "model_state_dict" in ck else ck
    n_user_cls, n_item_cls = len(ds.user_encoder.classes_), len(ds.item_encoder.classes_)
    if sd is not None:
        # embedding rows are addressed by encoder index: a mismatch means the encoders are not the training-time
        # ones and every lookup would silently hit the wrong row (the reference raises IndexError at best)
        for what, key, n_cls in (("user", "user_embedding.weight", n_user_cls), ("item", "item_embedding.weight", n_item_cls)):
            rows = int(sd[key].shape[0])
            if rows != n_cls:
                src = f"the pickled encoders in {enc_dir}" if enc_dir is not None else \
                    "encoders rebuilt from the interactions CSV / feature cache (pass --encoders DIR with the training-time pickles)"
                raise ValueError(f"checkpoint has {rows} {what} embedding rows but {src} define {n_cls} {what} ids")
    if n_tags is None:
        n_tags = int(sd["tag_embedding.weight"].shape[0]) if sd is not None else int(np.max(cache.tag)) + 1
    if len(cache.tag) and (int(np.max(cache.tag)) >= n_tags or int(np.min(cache.tag)) < 0):
        raise ValueError(f"feature cache holds tag index {int(np.max(cache.tag))} but the model has {n_tags} tag embeddings")
    model = FastMultimodalRecommender(
        n_users=n_user_cls, n_items=n_item_cls, n_tags=n_tags,
        num_numerical_features=cache.meta["num_numerical"], embedding_dim=m["embedding_dim"],
        vision_model_name=f"cached{cache.meta['vision_dim']}" if cache.meta["vision_dim"] else None,
        language_model_name=f"cached{cache.meta['language_dim']}" if cache.meta["language_dim"] else None,
        use_contrastive=False, num_attention_heads=m["num_attention_heads"], fusion_hidden_dims=list(m["fusion_hidden_dims"]),
        fusion_activation=m["fusion_activation"], use_batch_norm=m["use_batch_norm"],
        projection_hidden_dim=m["projection_hidden_dim"], final_activation=m["final_activation"], fusion_type=m["fusion_type"])
    if sd is not None:
        # strict: every parameter of the scoring path must come from the checkpoint (backbone / contrastive-head keys
        # are dropped by load_state_dict itself; BatchNorm's num_batches_tracked counters are not used)
        sd = {k: v for k, v in sd.items() if not k.startswith(_IGNORED_PREFIXES)}
        model.load_state_dict(sd, strict=True)
    dev = torch.device(device)
    missing_in_cache = [str(i) for i in ds.item_encoder.classes_ if str(i) not in cache.index]
    if missing_in_cache:
        print(f"warning: {len(missing_in_cache)} items of the item encoder have no row in the feature cache; "
              f"they score 0.0 like items without features in the reference (recommender.py:229-230)", file=sys.stderr)
    store = cache.to_store(dev, order=[str(i) for i in ds.item_encoder.classes_])
    item_range = None
    if shard is not None:                                   # (world, rank): contiguous item shard of this rank
        from .sharding import shard_range
        item_range = shard_range(n_item_cls, shard[0], shard[1])
    rec = FastRecommender(model, ds, dev, item_features=store, item_range=item_range)
    return rec, ds


# ----------------------------------------------------------------------------- generate
def select_users(args, all_users: Sequence[str]) -> List[str]:
    """``generate_recommendations.py:271-287``."""
    if args.users:
        return list(args.users)
    if args.user_file:
        return [ln.strip() for ln in Path(args.user_file).read_text().splitlines() if ln.strip()]
    if args.sample_users:
        import pandas as pd
        if len(all_users) < args.sample_users:
            return [str(u) for u in all_users]
        return pd.Series(list(all_users)).sample(n=args.sample_users, random_state=42).tolist()
    if getattr(args, "all_users", False):
        return [str(u) for u in all_users]
    return [str(u) for u in all_users[:5]]


def format_results(user_ids: Sequence[str], recs_per_user) -> Dict:
    """JSON shape of ``generate_recommendations.py:221-226``."""
    return {str(u): {"recommendations": [{"item_id": str(i), "score": float(s)} for i, s in recs]}
            for u, recs in zip(user_ids, recs_per_user)}


def cmd_generate(args) -> Dict:
    cfg = load_config(args.config)
    top_k = args.top_k or cfg["recommendation"]["top_k"]
    filter_seen = cfg["recommendation"]["filter_seen"] and not args.no_filter_seen
    rec, ds = build_recommender(cfg, args.checkpoint, args.cache, args.interactions, args.device, encoders=args.encoders)
    users = select_users(args, ds.user_encoder.classes_)
    if args.all_users or len(users) > 256:
        known = [u for u in users if u in rec.user_index]
        idx = np.fromiter((rec.user_index[u] for u in known), dtype=np.int64, count=len(known))
        s, i = rec.recommend_all(idx, top_k=top_k, filter_seen=filter_seen)
        s, i = s.cpu().numpy(), i.cpu().numpy()
        per_user = {u: [(rec.item_ids[int(b)], float(a)) for a, b in zip(s[r], i[r]) if b >= 0] for r, u in enumerate(known)}
        lists = [per_user.get(u, []) for u in users]
    else:
        lists = [rec.get_recommendations(u, top_k=top_k, filter_seen=filter_seen) for u in users]
    results = format_results(users, lists)
    out = Path(cfg["results_dir"]) / args.output if not Path(args.output).is_absolute() else Path(args.output)
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(results, indent=2))
    print(f"Saved recommendations for {len(users)} users to {out}")
    return results


# ----------------------------------------------------------------------------- evaluate
def cmd_evaluate(args) -> Dict:
    import pandas as pd
    from . import FullCatalogueEvaluator, RankingEvaluator, SampledRetrievalEvaluator
    cfg = load_config(args.config)
    top_k = args.top_k or cfg["recommendation"]["top_k"]
    world, rank, device = distributed_context(args.device)
    ranking = args.eval_task == "ranking"
    if world > 1 and (args.use_sampling or ranking):
        raise SystemExit("--use_sampling / --eval_task ranking score explicit pairs: run them on one GPU (no item-axis sharding)")
    rec, ds = build_recommender(cfg, args.checkpoint, args.cache, args.interactions, device,
                                shard=(world, rank) if world > 1 else None, encoders=args.encoders)
    test = pd.read_csv(args.test_data, dtype={"user_id": str, "item_id": str})
    ks = sorted(set([top_k] + [int(k) for k in (args.ks or [])]))
    sharded = None
    if world > 1:
        from .sharding import ShardedTopK
        # exact mode across shards: raw 64-slot lists are exchanged, the owning rank re-scores the merged candidates
        sharded = ShardedTopK(lambda users, k, fs: rec.recommend_all(users, top_k=k, filter_seen=fs, raw=True), rescore=rec.rescore)
    if ranking:                                             # evaluate.py:402-408: sampling is a retrieval-only switch
        ev = RankingEvaluator(rec, test, top_k=top_k, keep_predictions=bool(args.save_predictions))
    elif args.use_sampling:
        ev = SampledRetrievalEvaluator(rec, test, top_k=top_k, ks=ks, num_negatives=args.num_negatives,
                                       sampling_strategy=args.sampling_strategy, seed=args.seed,
                                       keep_predictions=bool(args.save_predictions))
    else:
        ev = FullCatalogueEvaluator(rec, test, top_k=top_k, ks=ks, filter_seen=cfg["recommendation"]["filter_seen"],
                                    keep_predictions=bool(args.save_predictions), sharded=sharded)
    results = ev.evaluate(novelty=True) if (args.novelty and not args.use_sampling and not ranking) else ev.evaluate()
    if rank != 0:                                           # every rank holds the same results; rank 0 writes them
        return results
    results_dir = Path(cfg["results_dir"])
    if args.save_predictions and "predictions" in results:          # evaluate.py:417-426
        preds = results.pop("predictions")
        p = results_dir / args.save_predictions
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(json.dumps({str(u): [{"item_id": str(i), "score": float(s)} for i, s in r] for u, r in preds.items()}, indent=2))
    results["evaluation_metadata"] = {"task": args.eval_task, "recommender_type": "fast_multimodal", "top_k": top_k,      # evaluate.py:429-434
                                      "test_file": args.test_data, "checkpoint_used": args.checkpoint}
    if "by_k" in results:
        results["by_k"] = {str(k): v for k, v in results["by_k"].items()}
    out = results_dir / args.output if not Path(args.output).is_absolute() else Path(args.output)
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(results, indent=2))
    print(json.dumps({k: v for k, v in results.items() if k.startswith("avg_") or k == "num_users_evaluated"}, indent=1))
    print(f"Evaluation results saved to {out}")
    return results


def make_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="pixelrec_multimodal_b200.cli", description=__doc__.split("\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)

    def common(p):
        p.add_argument("--config", type=str, default=None, help="reference-style YAML (model / recommendation sections)")
        p.add_argument("--checkpoint", type=str, default=None,
                       help="reference checkpoint (.pth with model_state_dict), or the directory holding it (then --checkpoint_name, "
                            "falling back to final_model.pth / last_model.pth like scripts/evaluate.py:76-90)")
        p.add_argument("--checkpoint_name", type=str, default="best_model.pth", help="scripts/evaluate.py:250")
        p.add_argument("--cache", type=str, required=True, help="packed feature cache directory (packed_cache.py)")
        p.add_argument("--interactions", "--train_data", dest="interactions", type=str, required=True,
                       help="train interactions CSV (user_id, item_id): histories / encoders (--train_data: the name scripts/evaluate.py uses)")
        p.add_argument("--encoders", type=str, default=None,
                       help="directory with the training-time user_encoder.pkl / item_encoder.pkl (default: searched next to the "
                            "checkpoint like scripts/evaluate.py; rebuilt from the tables only when none exist)")
        p.add_argument("--device", type=str, default="cuda:0")
        p.add_argument("--top_k", type=int, default=None)

    g = sub.add_parser("generate", help="scripts/generate_recommendations.py on the GPU path")
    common(g)
    g.add_argument("--users", type=str, nargs="+")
    g.add_argument("--user_file", type=str)
    g.add_argument("--sample_users", type=int)
    g.add_argument("--all_users", action="store_true", help="every user of the interaction table (batched API)")
    g.add_argument("--no_filter_seen", action="store_true")
    g.add_argument("--use_diversity", action="store_true",
                   help="accepted like scripts/generate_recommendations.py:252; the reference Recommender has no diversity-aware method, "
                        "so (as there, :206-212) a warning is printed and the standard lists are produced")
    g.add_argument("--output", type=str, default="recommendations.json")
    g.set_defaults(fn=cmd_generate)

    e = sub.add_parser("evaluate", help="scripts/evaluate.py (retrieval and ranking tasks) on the GPU path")
    common(e)
    e.add_argument("--test_data", type=str, required=True)
    e.add_argument("--eval_task", type=str, default="retrieval", choices=["retrieval", "ranking"],
                   help="retrieval: top-K lists against the test positives; ranking: order of each user's own test items (tasks.py:776-901)")
    e.add_argument("--use_sampling", action="store_true", help="positives + sampled negatives (the reference script's default protocol; "
                                                               "here the full catalogue is the default -- it is what this path is for)")
    e.add_argument("--no_sampling", dest="use_sampling", action="store_false", help="scripts/evaluate.py:247 (the default here)")
    e.add_argument("--num_negatives", type=int, default=100, help="negatives per user (scripts/evaluate.py:248 defaults to 20)")
    e.add_argument("--recommender_type", type=str, default="multimodal",
                   help="scripts/evaluate.py:239-241; only the multimodal recommender lives on this path (the baselines stay in the reference)")
    e.add_argument("--num_workers", type=int, default=1,
                   help="accepted for compatibility (scripts/evaluate.py:245); the evaluation is one batched GPU pass, forked workers "
                        "cannot share a CUDA context: values > 1 are ignored with a warning")
    e.add_argument("--warmup_recommender_cache", action="store_true",
                   help="accepted for compatibility (scripts/evaluate.py:244); the item records are resident after the one precompute")
    e.add_argument("--sampling_strategy", type=str, default="random", choices=["random", "popularity", "popularity_inverse"],
                   help="how the negatives are drawn (evaluate.py / tasks.py:221-308)")
    e.add_argument("--seed", type=int, default=20261018)
    e.add_argument("--ks", type=int, nargs="*", help="extra cut-offs reported under by_k")
    e.add_argument("--save_predictions", type=str, default=None)
    e.add_argument("--novelty", action="store_true", help="add the novelty / coverage / personalization block (tasks.py:637-714)")
    e.add_argument("--output", type=str, default="evaluation_results.json")
    e.set_defaults(fn=cmd_evaluate)
    return ap


def resolve_compat(args):
    """The reference scripts' flags that have no work to do on this path, and the checkpoint-directory form."""
    import logging
    if getattr(args, "recommender_type", "multimodal") not in ("multimodal", "fast_multimodal"):
        raise SystemExit(f"--recommender_type {args.recommender_type}: only the multimodal recommender runs on the GPU path "
                         "(random / popularity / item_knn / user_knn baselines: use the reference's scripts/evaluate.py)")
    if getattr(args, "num_workers", 1) > 1:
        logging.warning("--num_workers %d ignored: the evaluation is one batched pass on the GPU (no forked workers)", args.num_workers)
        args.num_workers = 1
    if getattr(args, "use_diversity", False):
        print("Warning: Diversity method not implemented. Falling back to standard recommendations.")
    if args.checkpoint and Path(args.checkpoint).is_dir():
        d = Path(args.checkpoint)
        for name in [args.checkpoint_name, "best_model.pth", "final_model.pth", "last_model.pth"]:
            if (d / name).is_file():
                args.checkpoint = str(d / name)
                break
        else:
            raise SystemExit(f"no checkpoint ({args.checkpoint_name}, best_model.pth, final_model.pth, last_model.pth) in {d}")
    return args


def main(argv: Optional[Sequence[str]] = None):
    args = resolve_compat(make_parser().parse_args(argv))
    return args.fn(args)


if __name__ == "__main__":
    main(sys.argv[1:])

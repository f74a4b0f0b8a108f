#!/usr/bin/env python
"""Tiny invocation of every kernel family of libpxr.so, meant to run under `compute-sanitizer --tool memcheck` (and racecheck)
on a B200: item precompute (tensor-pipe and SIMT), fused scoring for the three fusions (raw + exact mode, item splits, seen
filter), generic top-K path, explicit pairs, merge, metrics, sampler, novelty, Gini, intra-list similarity.

  compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_smoke.py
(compute-sanitizer is closed on the round-2 GPU pool; run plainly the script is a crash / launch-error check of every
kernel family on edge shapes: unaligned and short metric lists, partial groups, paged top-K, forced small-batch shape.)
"""
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from pixelrec_multimodal_b200 import synthetic as syn                                   # noqa: E402
from pixelrec_multimodal_b200.engine import merge_topk, ranking_metric_sums, sample_candidates, weighted_candidates   # noqa: E402
from pixelrec_multimodal_b200.evaluation import beyond_accuracy_metrics, gini_coefficient, intra_list_similarity, novelty_tables  # noqa: E402
from tests import _cases as cs                                                            # noqa: E402


def main():
    dev = torch.device("cuda:0")
    # `--simt-only`: leave the tcgen05 / cluster kernels out (the sanitizer run of this round used it: memcheck on the
    # generic, merge, metric, sampler and diversity kernels; the fused kernels are covered by the parity tests)
    paths = ("simt",) if "--simt-only" in sys.argv else ("auto", "simt")
    for fusion in ("gated", "concatenate", "attention"):
        for path in paths:
            spec = syn.ModelSpec(n_users=70, n_items=733, fusion_type=fusion, fusion_activation="relu" if fusion == "attention" else "silu")
            sd = syn.make_state_dict(spec, seed=3)
            feats = syn.make_item_features(spec, seed=3)
            indptr, idx, _ = syn.make_histories(spec.n_users, spec.n_items, seed=3, lo=3, hi=20)
            model = cs.torch_model_from(spec, sd, kernel_path=path)
            eng = model.engine("catalogue")
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            eng.precompute_items(model.item_embedding.weight.detach(), t(feats["tag_idx"]), t(feats["vis"]), t(feats["txt"]), t(feats["num"]))
            uemb = model.user_embedding.weight.detach()
            users = torch.arange(spec.n_users, device=dev)
            for resc in (True, False):
                eng.set_rescore(resc)
                s, i = eng.score_topk(uemb, users, 50, t(indptr), t(idx))
                s2, i2 = eng.score_topk(uemb, users[:3], 7)
            s3, i3 = eng.score_topk(uemb, users[:5], 100, t(indptr[:6]), t(idx))          # top_k > 64: paged fused passes / generic path
            if eng.active_path == "tcgen05" and fusion != "attention":                      # small-batch tile shape (forced), 1 and 3 users
                eng.set_small_batch(1)
                for nu in (1, 3):
                    eng.score_topk(uemb, users[:nu], 50, t(indptr[:nu + 1]), t(idx))
                eng.set_small_batch(-1)
            eng.score_pairs(uemb, users[:50], torch.arange(50, device=dev), want_logit=True)
            eng.rescore_topk(uemb, users[:9], torch.randint(-1, spec.n_items, (9, 64), device=dev, dtype=torch.int32), 20)
            torch.cuda.synchronize()
            print("ok", fusion, path, eng.active_path, flush=True)
    sc = torch.rand((5, 40, 50), device=dev).sort(dim=2, descending=True).values
    ix = torch.randint(0, 1000, (5, 40, 50), device=dev, dtype=torch.int32)
    merge_topk(sc, ix); merge_topk(sc[:, :, :7].contiguous(), ix[:, :, :7].contiguous())
    sc100 = torch.rand((3, 9, 100), device=dev).sort(dim=2, descending=True).values
    merge_topk(sc100, torch.randint(0, 1000, (3, 9, 100), device=dev, dtype=torch.int32))
    topk = torch.randint(-1, 300, (500, 50), device=dev, dtype=torch.int32)
    gt_ptr = torch.arange(501, device=dev, dtype=torch.int64) * 2
    gt_idx = torch.randint(0, 300, (1000,), device=dev, dtype=torch.int32)
    ranking_metric_sums(topk, gt_ptr, gt_idx, [10, 50], recall_den=torch.full((500,), 3, dtype=torch.int32))
    ranking_metric_sums(torch.randint(-1, 300, (70, 100), device=dev, dtype=torch.int32), gt_ptr[:71], gt_idx[:140], [10, 100])
    # K5 warp kernel: unaligned list base (scalar loads), short / full-width lists, a partial last group, many cut-offs
    ranking_metric_sums(topk[1:], gt_ptr[1:] - 2, gt_idx[2:], [10, 50])
    for kk in (7, 13, 64):
        tk = torch.randint(-1, 300, (333, kk), device=dev, dtype=torch.int32)
        ranking_metric_sums(tk, gt_ptr[:334], gt_idx[:666], [1, 3, 5, 7] if kk == 7 else [5, kk])
    wts = torch.rand(300, dtype=torch.float64, device=dev) + 0.01
    pos_ptr = torch.arange(41, dtype=torch.int64) * 2
    pos_idx = torch.sort(torch.randint(0, 300, (40, 2)), dim=1).values.reshape(-1).to(torch.int32).to(dev)
    sample_candidates(torch.arange(40), pos_ptr, pos_idx, 300, 20, 7)
    weighted_candidates(torch.arange(40, device=dev), pos_ptr.to(dev), pos_idx, wts, 20, seed=7, stride=32)
    si, iif, n_pop = novelty_tables(np.random.default_rng(0).integers(0, 300, 2000), 100, 300)
    beyond_accuracy_metrics(topk, si, iif, n_pop)
    gini_coefficient(topk, 300, include_zero=True)
    intra_list_similarity(topk, embeddings=torch.randn(300, 320, device=dev))
    torch.cuda.synchronize()
    print("sanitize smoke OK")


if __name__ == "__main__":
    main()

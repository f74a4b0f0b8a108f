#!/usr/bin/env python
"""Achieved HBM GB/s of the memory-bound stages next to the fused scoring kernel:
K1+K2 item precompute, K4 S-way top-K merge, K5 ranking metrics.  Algorithmic bytes per unit are the
§8(d) figures (stated in DESIGN.md §5); time = CUDA events around `reps` replays of a CUDA graph of the stage's call
(device time without host gaps: several of these stages are shorter than the Python cost of their call), after a
warm-up, L2 flushed before the timed group; every stage's working set is larger than the L2.  One JSON line per stage.

  python scripts/bench_stages.py [--fusion gated] [--items 96282]
"""
import argparse, json, sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from pixelrec_multimodal_b200 import FastMultimodalRecommender, synthetic as syn   # noqa: E402
from pixelrec_multimodal_b200.engine import merge_topk, ranking_metric_sums, sample_candidates, weighted_candidates   # noqa: E402


EAGER = False      # --eager: time the Python calls back to back instead of graph replays (includes host gaps)


def timed(fn, reps, flush):
    # three untimed calls: the first two of a stage that allocates its workspace grow the caching allocator's pool
    # (cudaMalloc on the host path: 11 ms and 5 ms for the 222 MB concat workspace, profiles/r01_diag_precompute_concat.log)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    run = fn
    if not EAGER:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            run = g.replay
            run(); torch.cuda.synchronize()
        except Exception as ex:                                   # a stage that cannot be captured is timed eagerly
            print(json.dumps({"note": f"graph capture failed, eager timing: {type(ex).__name__}: {str(ex)[:120]}"}), flush=True)
            torch.cuda.synchronize()
            run = fn
    flush.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fusion", default="gated")
    ap.add_argument("--items", type=int, default=96282)
    ap.add_argument("--users", type=int, default=1 << 20)
    ap.add_argument("--eager", action="store_true", help="time the Python calls (with their host gaps) instead of graph replays")
    args = ap.parse_args()
    global EAGER
    EAGER = args.eager
    peak = json.loads((REPO / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (REPO / "MEASURED_PEAKS.json").exists() else 6650.0
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    spec = syn.ModelSpec(n_users=4096, n_items=args.items, fusion_type=args.fusion)
    sd, feats, _ = syn.torch_workload(spec, dev, seed=1, with_histories=False)
    m = FastMultimodalRecommender(n_users=spec.n_users, n_items=spec.n_items, n_tags=spec.n_tags, num_numerical_features=7,
                                  embedding_dim=64, vision_model_name="cached512", language_model_name="cached384",
                                  fusion_type=args.fusion).to(dev)
    m.load_state_dict(sd, strict=False)
    e = m.engine("catalogue")
    D, Dv, Dl, F, M = 64, 512, 384, 7, 6
    out = []
    # ---- K1 + K2: gather + projections -> item records (the fast-path extras of the fusion type are part of the stage)
    ms = timed(lambda: e.precompute_items(m.item_embedding.weight.detach(), feats["tag_idx"], feats["vis"], feats["txt"], feats["num"],
                                          validate=False), 3, flush)
    extra = {"gated": 8 * 4, "concatenate": 512 * 2, "attention": 1032 * 16}[args.fusion] if e.active_path == "tcgen05" else 0
    b_item = 4 * (2 * D + Dv + Dl + F) + 8 + 4 * (M - 1) * D + extra      # read features/embeddings/tag index + write record
    out.append(dict(stage="K1+K2 item precompute", fusion=args.fusion, units=args.items, unit="items", ms=ms,
                    bytes_per_unit=b_item, achieved_gbs=args.items * b_item / ms / 1e6, peak_gbs=peak))
    # ---- K4: S-way merge of per-shard top-K lists
    S, K, n = 8, 50, args.users // 8
    sc = torch.rand((S, n, K), device=dev).sort(dim=2, descending=True).values
    ix = torch.randint(0, 1 << 20, (S, n, K), device=dev, dtype=torch.int32)
    ms = timed(lambda: merge_topk(sc, ix), 5, flush)
    b = 8 * K * S + 8 * K
    out.append(dict(stage="K4 top-K merge", S=S, K=K, units=n, unit="users", ms=ms, bytes_per_unit=b,
                    achieved_gbs=n * b / ms / 1e6, peak_gbs=peak))
    # ---- K5: ranking metrics @10 / @50 (one positive per user, leave-one-out)
    n = args.users
    topk = torch.randint(0, args.items, (n, K), device=dev, dtype=torch.int32)
    gt_indptr = torch.arange(n + 1, device=dev, dtype=torch.int64)
    gt_idx = torch.randint(0, args.items, (n,), device=dev, dtype=torch.int32)
    ms = timed(lambda: ranking_metric_sums(topk, gt_indptr, gt_idx, [10, 50], as_device=True), 5, flush)
    b = 4 * K + 8 + 4
    out.append(dict(stage="K5 metrics @10/@50", K=K, units=n, unit="users", ms=ms, bytes_per_unit=b,
                    achieved_gbs=n * b / ms / 1e6, peak_gbs=peak, note="pxr_metrics (two kernels), result left on the device"))
    # the same with a hit for every fifth user (a hit costs the float64 arithmetic; users without one add exact zeros)
    gt_hit = gt_idx.clone()
    sel = torch.arange(0, n, 5, device=dev)
    gt_hit[sel] = topk[sel, torch.randint(0, K, (sel.numel(),), device=dev)]
    ms = timed(lambda: ranking_metric_sums(topk, gt_indptr, gt_hit, [10, 50], as_device=True), 5, flush)
    out.append(dict(stage="K5 metrics @10/@50, 20 % of users with a hit", K=K, units=n, unit="users", ms=ms, bytes_per_unit=b,
                    achieved_gbs=n * b / ms / 1e6, peak_gbs=peak))
    for o in out:
        o["frac"] = o["achieved_gbs"] / o["peak_gbs"]
        print(json.dumps(o), flush=True)
    # ---- K3r: the fp32 re-score of exact mode = pxr_score_pairs over 64 candidates per user of a 4 096-user block
    # (CUDA-core fp32 arithmetic: reported as pairs/s and fp32 TFLOP/s, not against the HBM roof)
    npairs = 4096 * 64
    uu = torch.arange(4096, device=dev).repeat_interleave(64)
    ii = torch.randint(0, args.items, (npairs,), device=dev)
    uemb = m.user_embedding.weight.detach()
    ms = timed(lambda: e.score_pairs(uemb, uu, ii), 5, flush)
    flop = 2 * (64 * 512 + 512 * 256 + 256 * 128 + 128) if args.fusion != "concatenate" else 2 * (384 * 512 + 512 * 256 + 256 * 128 + 128)
    print(json.dumps(dict(stage="K3r fp32 re-score (pxr_score_pairs)", fusion=args.fusion, units=npairs, unit="pairs", ms=ms,
                          pairs_per_s=npairs / ms * 1e3, fp32_tflops=npairs * flop / ms / 1e9)), flush=True)
    # ---- K6: candidate samplers of the sampled protocol, 100 negatives + 1 positive per user (compute-bound: per (user, item) a
    # hash and a float64 log for the weighted one; reported as users/s and (user, item) keys/s)
    nu = 8192
    pos_ptr = torch.arange(nu + 1, device=dev, dtype=torch.int64)
    pos = torch.randint(0, args.items, (nu,), device=dev, dtype=torch.int32)
    us = torch.arange(nu, device=dev, dtype=torch.int64)
    w = torch.from_numpy(np.random.default_rng(0).zipf(1.5, args.items).clip(max=1000).astype(np.float64)).to(dev)
    ms = timed(lambda: sample_candidates(us, pos_ptr, pos, args.items, 100, seed=1, stride=101), 5, flush)
    print(json.dumps(dict(stage="K6 uniform sampler (pxr_sample_candidates)", units=nu, unit="users", items=args.items, ms=ms,
                          users_per_s=nu / ms * 1e3)), flush=True)
    ms = timed(lambda: weighted_candidates(us, pos_ptr, pos, w, 100, seed=1, stride=101), 3, flush)
    print(json.dumps(dict(stage="K6 weighted sampler (pxr_weighted_candidates)", units=nu, unit="users", items=args.items, ms=ms,
                          users_per_s=nu / ms * 1e3, keys_per_s=nu * args.items / ms * 1e3)), flush=True)


if __name__ == "__main__":
    main()

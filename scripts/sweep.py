#!/usr/bin/env python
"""BASELINE.json configs[4]: catalogue-size x user-batch sweep of the fused scoring + top-50 kernel on one GPU,
each point against the bf16 tensor roofline (same definitions as bench.py).  One JSON line per point.

  python scripts/sweep.py [--fusion gated concatenate] [--items 10000 100000 1000000] [--batch 1 64 1024 8192]
"""
import argparse, json, sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench                                                                     # noqa: E402
from pixelrec_multimodal_b200 import FastMultimodalRecommender, synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fusion", nargs="+", default=["gated", "concatenate"])
    ap.add_argument("--items", nargs="+", type=int, default=[10_000, 100_000, 1_000_000])
    ap.add_argument("--batch", nargs="+", type=int, default=[1, 64, 1024, 8192])
    ap.add_argument("--min-ms", type=float, default=300.0, help="repeat launches until this much kernel time is accumulated")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    pk = bench.peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    NU = max(args.batch)
    for fusion in args.fusion:
        for NI in args.items:
            spec = syn.ModelSpec(n_users=NU, n_items=NI, fusion_type=fusion)
            sd, feats, hist = syn.torch_workload(spec, dev, seed=11)
            syn.condition_like_trained(sd, spec, feats)
            m = FastMultimodalRecommender(n_users=NU, n_items=NI, n_tags=spec.n_tags, num_numerical_features=7, embedding_dim=64,
                                          vision_model_name="cached512", language_model_name="cached384", fusion_type=fusion).to(dev)
            m.load_state_dict(sd, strict=False)
            e = m.engine("catalogue")
            e.precompute_items(m.item_embedding.weight.detach(), feats["tag_idx"], feats["vis"], feats["txt"], feats["num"])
            del feats
            uemb = m.user_embedding.weight.detach()
            wp = bench.w_pair(fusion, 64, [512, 256, 128])
            for B in args.batch:
                users = torch.arange(B, device=dev)
                ip, ix = hist["train_indptr"][:B + 1], hist["train_idx"]
                for _ in range(3):
                    e.score_topk(uemb, users, 50, ip, ix)
                torch.cuda.synchronize()
                flush.zero_()
                e.profile(True)
                reps, est = 0, 0.0
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                while reps < 3 or (est < args.min_ms and reps < 200):
                    e.score_topk(uemb, users, 50, ip, ix)
                    reps += 1
                    if reps % 3 == 0:
                        e1.record(); torch.cuda.synchronize(); est = e0.elapsed_time(e1)
                e1.record(); torch.cuda.synchronize()
                k_ms, k_n = e.profile_read()
                e.profile(False)
                ms = e0.elapsed_time(e1) / reps
                pairs = B * NI
                tf = pairs * wp / (k_ms / k_n * 1e-3) / 1e12
                print(json.dumps({"fusion": fusion, "n_items": NI, "user_batch": B, "path": e.active_path, "ms_per_call": ms,
                                  "kernel_ms": k_ms / k_n, "pairs_per_s": pairs / (ms * 1e-3), "users_per_s": B / (ms * 1e-3),
                                  "tflops": tf, "frac_of_bf16_peak": tf / pk["tf_sustained"], "reps": reps}), flush=True)
            del e, m, sd, hist
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

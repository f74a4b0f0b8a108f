#!/usr/bin/env python
"""BASELINE.json configs[4]: catalogue-size x user-batch sweep of the fused scoring + top-50 kernel on one GPU,
each point against the bf16 tensor roofline (same definitions as bench.py).  One JSON line per point.

  python scripts/sweep.py [--fusion gated concatenate attention] [--items 10000 100000 1000000 10000000]
                          [--batch 1 64 1024 8192] [--dims 64 128 256 512]
  torchrun --nproc-per-node N scripts/sweep.py ...      item-axis shards + the owned all-to-all exchange (bench.run_config)

Embedding dims other than 64 run fused for concat fusion (layer 1 is applied as partials); gated / attention then take
the generic fp32 kernels (reported with their path).  The 16.5 KB attention records bound that fusion to ~1 M items per GPU.
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
    ap.add_argument("--dims", nargs="+", type=int, default=[64])
    ap.add_argument("--activation", nargs="+", default=["relu"], help="fusion_activation of the model (one-GPU mode)")
    ap.add_argument("--top-k", nargs="+", type=int, default=[50], help="list lengths; > 64 runs one fused pass per 64-slot page (one-GPU mode)")
    ap.add_argument("--small-batch", type=int, default=-1, choices=[-1, 0, 1], help="pxr_set_small_batch mode (one-GPU mode): -1 auto, 0 plain tile shape, 1 forced")
    ap.add_argument("--min-ms", type=float, default=300.0, help="repeat launches until this much kernel time is accumulated")
    args = ap.parse_args()
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return sharded(args)
    dev = torch.device("cuda:0")
    pk = bench.peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    NU = max(args.batch)
    for fusion, NI, Dm, act in [(f, n, d, a) for f in args.fusion for n in args.items for d in args.dims for a in args.activation]:
        if True:
            if fusion == "attention" and NI > 2_000_000:
                continue                                   # 16.5 KB of MMA fragments per item
            spec = syn.ModelSpec(n_users=NU, n_items=NI, fusion_type=fusion, embedding_dim=Dm, fusion_activation=act)
            sd, feats, hist = syn.torch_workload(spec, dev, seed=11)
            syn.condition_like_trained(sd, spec, feats)
            m = FastMultimodalRecommender(n_users=NU, n_items=NI, n_tags=spec.n_tags, num_numerical_features=7, embedding_dim=Dm,
                                          vision_model_name="cached512", language_model_name="cached384", fusion_type=fusion,
                                          fusion_activation=act).to(dev)
            m.load_state_dict(sd, strict=False)
            e = m.engine("catalogue")
            e.set_small_batch(args.small_batch)
            e.precompute_items(m.item_embedding.weight.detach(), feats["tag_idx"], feats["vis"], feats["txt"], feats["num"])
            del feats
            uemb = m.user_embedding.weight.detach()
            wp = bench.w_pair(fusion, Dm, [512, 256, 128])
            slow = e.active_path != "tcgen05"
            for B, top_k in [(b, k) for b in args.batch for k in args.top_k]:
                if slow and B * NI > 3e8:
                    continue                               # generic fp32 kernels: ~40 M pairs/s
                users = torch.arange(B, device=dev)
                ip, ix = hist["train_indptr"][:B + 1], hist["train_idx"]
                huge = B * NI > 1.5e10                     # a 10 M-item catalogue x 8 192 users is ~30 s per call
                for _ in range(1 if huge else 3):
                    e.score_topk(uemb, users, top_k, ip, ix)
                torch.cuda.synchronize()
                flush.zero_()
                e.profile(True)
                reps, est = 0, 0.0
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                while reps < (1 if huge else 3) or (est < args.min_ms and reps < 200):
                    e.score_topk(uemb, users, top_k, ip, ix)
                    reps += 1
                    if huge or reps % 3 == 0:
                        e1.record(); torch.cuda.synchronize(); est = e0.elapsed_time(e1)
                e1.record(); torch.cuda.synchronize()
                k_ms, k_n = e.profile_read()
                e.profile(False)
                ms = e0.elapsed_time(e1) / reps
                pairs = B * NI
                tf = pairs * wp / (k_ms / k_n * 1e-3) / 1e12
                print(json.dumps({"fusion": fusion, "n_items": NI, "embedding_dim": Dm, "user_batch": B, "activation": act, "top_k": top_k, "small_batch_mode": args.small_batch, "path": e.active_path, "exact_rescore": bool(e.rescore), "ms_per_call": ms,
                                  "kernel_ms": k_ms / k_n, "pairs_per_s": pairs / (ms * 1e-3), "users_per_s": B / (ms * 1e-3),
                                  "tflops": tf, "frac_of_bf16_peak": tf / pk["tf_sustained"], "reps": reps}), flush=True)
            del e, m, sd, hist
            torch.cuda.empty_cache()


def sharded(args):
    """N ranks under torchrun: every point through bench.run_config (item shards, overlapped owned exchange)."""
    import os
    import torch.distributed as dist
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    ns = argparse.Namespace(seed=11, path="auto", user_block=0, cpu_seconds=0.0, literal_seconds=0.0)
    for fusion, NI, Dm, B in [(f, n, d, b) for f in args.fusion for n in args.items for d in args.dims for b in args.batch]:
        if fusion == "attention" and NI // world > 2_000_000:
            continue
        name = f"E:{fusion}:{NI}:{Dm}:{B}"
        users = max(B * world * 8, B * world)
        bench.CONFIGS[name] = (users, NI, fusion, f"configs[4] sweep point: {fusion}, {NI} items, embedding_dim {Dm}, {B} users per GPU and step")
        bench.CONFIG_DIMS[name] = Dm
        ns.user_block = B
        steps = 3 if B * (NI // world) > 2e8 else 10
        r = bench.run_config(ns, name, steps, 3, world, rank, dev, shard="items", with_e2e=False, with_cpu=False, with_checks=False)
        if rank == 0:
            print(json.dumps({"fusion": fusion, "n_items": NI, "embedding_dim": Dm, "user_batch_per_gpu": B, "n_gpus": world,
                              "path": r["run"]["kernel_path"], "pairs_per_s": r["value"], "users_per_s": r["users_per_sec"],
                              "ms_per_step": r["ms_per_step"], "kernel_ms": r["roofline"]["kernel_ms_avg"],
                              "frac_of_bf16_peak": r["roofline"]["frac"], "kernel_share_of_step": r["roofline"]["kernel_share_of_step"]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

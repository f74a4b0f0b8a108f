#!/usr/bin/env python
"""bench.py — headline benchmark of the full-catalogue scoring + top-K path.

A "step" is one pass of the hot path over one block of users: every user of the
block scored against EVERY item of the catalogue, seen items filtered, top-50
kept (pxr_score_topk; with N>1 GPUs each rank owns a contiguous item shard and
the per-shard lists are merged after one NCCL all-gather).

Workload at N=1 = BASELINE.json configs[1]: gated fusion, CLIP-512 + SBERT-384
cached features, Pixel200K-shaped synthetic (200 000 users x 96 282 items), top-50.
Per-GPU work is fixed as N grows (users per step = user_block * N, items per
rank = NI / N): "scaling": "weak".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

`--impl reference` times the reference's CPU path (the numpy oracle port: the
reference is pure Python/PyTorch and cannot travel to the GPU box) on the
host cores for the same metric and config.  The oracle is only ever the CPU
arm here; the product path never touches it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

CONFIGS = {
    # name: (n_users, n_items, fusion, description)
    "A": (1_000, 2_000, "concatenate", "configs[0] simple_config_example concat 1K x 2K"),
    "B": (200_000, 96_282, "gated", "configs[1] gated, CLIP-512 + SBERT-384, Pixel200K-shaped 200K x 96K, top-50"),
    "Bc": (200_000, 96_282, "concatenate", "concat fusion at the configs[1] shape (200K x 96K), top-50 (not a BASELINE config: kernel comparison)"),
    "C": (1_001_822, 100_541, "attention", "configs[2] attention + numerical, Pixel1M-shaped 1M x 100K"),
    "D": (8_886_078, 407_082, "gated", "configs[3] Pixel8M-shaped 8.9M x 407K item-sharded"),
}
TOP_K = 50
METRIC = "scored user-item pairs/sec (full catalogue, top-50 per user)"
UNIT = "pairs/s"


def w_pair(fusion: str, D: int, H):
    """Algorithmic tensor FLOPs per scored pair: the MLP tail
    2*(sum H_{l-1} H_l + H_L), plus the per-pair first layer 2*D*H1 when the
    fused vector depends on the pair (gated / attention)."""
    tail = 2 * (sum(a * b for a, b in zip(H[:-1], H[1:])) + H[-1])
    return tail + (0 if fusion == "concatenate" else 2 * D * H[0])


def ncu_traffic(fusion: str, users_per_launch: int, items_per_rank: int):
    """DRAM bytes (read + write) of ONE launch of the fused kernel from the committed `ncu --set full` capture of this
    same workload shape (profiles/ncu_traffic.json, written by scripts/ncu_summary.py traffic); None when no capture matches."""
    p = REPO / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return None
    for e in json.loads(p.read_text()).get("captures", []):
        if e["fusion"] == fusion and e["users_per_launch"] == users_per_launch and e["items_per_rank"] == items_per_rank:
            return e["dram_bytes_per_launch"]
    return None


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm_gbs=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sustained=j["bf16_tflops_sustained"], src="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[x.strip() for x in ln.split(",")] for ln in out.splitlines() if ln.strip()]
        isnum = lambda v: v.replace(".", "", 1).isdigit()
        sm = [float(r[0]) for r in rows if len(r) >= 7 and isnum(r[0])]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and isnum(r[1])]
        pw = [float(r[2]) for r in rows if len(r) >= 7 and isnum(r[2])]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(rows)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port (batched reference forward + stable top-K)
# ---------------------------------------------------------------------------
_CPU_WL = {}


def cpu_workload(cfg_name: str, seed: int):
    if (cfg_name, seed) not in _CPU_WL:
        import torch
        from pixelrec_multimodal_b200 import synthetic as syn
        from tests import _cases as cs
        NU, NI, fusion, _ = CONFIGS[cfg_name]
        spec = syn.ModelSpec(n_users=NU, n_items=NI, fusion_type=fusion)
        sd, feats, hist = syn.torch_workload(spec, "cpu", seed=seed)
        syn.condition_like_trained(sd, spec, feats)
        torch.set_num_threads(os.cpu_count() or 1)
        _CPU_WL[(cfg_name, seed)] = dict(sd=sd, feats=feats, indptr=hist["train_indptr"], idx=hist["train_idx"],
                                         cfg=cs.spec_cfg(spec), NU=NU, NI=NI, fusion=fusion)
    return _CPU_WL[(cfg_name, seed)]


def cpu_arm(cfg_name: str, seconds: float, seed: int, first_user: int = 0):
    """Times the reference's CPU path (oracle/pxr_oracle_torch.py: the reference
    forward restated op-for-op in PyTorch eager fp32, pinned to the reference's
    outputs in tests/) on a bounded sample of the SAME workload: `n` users x the
    full catalogue as flattened pairs (multimodal.py:528-610), seen items
    dropped, stable top-50 (recommender.py:88-106).  Returns pairs/s."""
    import torch
    from oracle import pxr_oracle_torch as ot
    wl = cpu_workload(cfg_name, seed)
    NU, NI, fusion = wl["NU"], wl["NI"], wl["fusion"]
    threads = torch.get_num_threads()

    def run(users):
        ot.recommend_block(wl["sd"], wl["cfg"], torch.as_tensor(users), wl["feats"], TOP_K, wl["indptr"], wl["idx"])

    t0 = time.perf_counter(); run(np.arange(first_user, first_user + 2) % NU); t1 = (time.perf_counter() - t0) / 2
    n = int(max(2, min(NU, seconds / max(t1, 1e-3))))
    users = np.arange(first_user + 2, first_user + 2 + n) % NU
    t0 = time.perf_counter(); run(users); dt = time.perf_counter() - t0
    return dict(value=n * NI / dt, unit=UNIT, cores=threads, kind="port", seconds=dt, users=n,
                sample=f"{n} users x {NI} items of config {cfg_name} ({fusion}): PyTorch-eager fp32 restatement of the "
                       f"reference forward on flattened pairs ({threads} intra-op threads of {os.cpu_count()} cores) "
                       f"+ seen filter + stable top-{TOP_K}")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    for s in range(args.warmup + args.steps):
        r = cpu_arm(args.config, per_step, args.seed, first_user=1000 * s)
        if s >= args.warmup:
            vals.append(r)
    tot_pairs = sum(r["users"] for r in vals) * CONFIGS[args.config][1]
    tot_s = sum(r["seconds"] for r in vals)
    v = tot_pairs / tot_s
    NU, NI, fusion, desc = CONFIGS[args.config]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, len(vals)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "users_per_sec": v / NI,
            "config": {"workload": desc, "n_users": NU, "n_items": NI, "fusion": fusion, "top_k": TOP_K},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": vals[-1]["cores"], "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def b200_arm(args):
    import torch
    import torch.distributed as dist
    from pixelrec_multimodal_b200 import FastMultimodalRecommender, FastRecommender, ItemFeatureStore, synthetic as syn
    from pixelrec_multimodal_b200.engine import merge_topk
    from pixelrec_multimodal_b200.sharding import allgather_topk, allgather_topk_finish, allgather_topk_start, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    NU, NI, fusion, desc = CONFIGS[args.config]
    spec = syn.ModelSpec(n_users=NU, n_items=NI, fusion_type=fusion)
    sd, feats, hist = syn.torch_workload(spec, dev, seed=args.seed)      # identical on every rank (same seed)
    # BatchNorm statistics matched to the activations and logits spread to std 2, like a trained checkpoint
    syn.condition_like_trained(sd, spec, feats)

    model = FastMultimodalRecommender(
        n_users=NU, n_items=NI, n_tags=spec.n_tags, num_numerical_features=spec.num_numerical_features,
        embedding_dim=spec.embedding_dim, vision_model_name=f"cached{spec.vision_dim}",
        language_model_name=f"cached{spec.language_dim}", use_contrastive=False, fusion_type=fusion,
        fusion_hidden_dims=list(spec.fusion_hidden_dims), kernel_path=args.path).to(dev)
    model.load_state_dict(sd, strict=False)

    class _Enc:
        def __init__(self, n, p): self.classes_ = _LazyIds(n, p)

    class _DS:
        user_encoder, item_encoder, interactions = _Enc(NU, "u"), _Enc(NI, "i"), None

    lo, hi = shard_range(NI, world, rank)
    store = ItemFeatureStore(feats["tag_idx"], feats["vis"], feats["txt"], feats["num"])
    rec = FastRecommender(model, _DS(), dev, item_features=store, n_users=NU, n_items=NI,
                          history=(hist["train_indptr"], hist["train_idx"]), item_range=(lo, hi),
                          user_block=args.user_block * world)
    eng = rec.engine()
    torch.cuda.synchronize()

    B = args.user_block * world                                   # users per step (whole job)
    n_blocks = max(1, NU // B)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    uemb = model.user_embedding.weight.detach()
    d_indptr, d_idx = rec.device_history()
    merges = [0]

    all_users = torch.arange(NU, device=dev)

    def score_block(s):
        """inputs already in HBM: the block's user indices and the resident history CSR"""
        b = s % n_blocks
        u0 = b * B
        n = min(NU, u0 + B) - u0
        sc, ix = eng.score_topk(uemb, all_users[u0:u0 + n], TOP_K, d_indptr[u0:u0 + n + 1], d_idx)
        return sc, ix, n

    def run_steps(first, count):
        """`count` steps.  N > 1: the all-gather of step s (one packed collective on NCCL's stream) overlaps the
        scoring kernel of step s + 1; its merge is enqueued behind that kernel."""
        users = 0
        pending = None
        for s in range(first, first + count):
            sc, ix, n = score_block(s)
            users += n
            if world > 1:
                handle = allgather_topk_start(sc, ix)
                if pending is not None:
                    merge_topk(*allgather_topk_finish(pending)); merges[0] += 1
                pending = handle
            flush.zero_()                                         # L2 flush between steps (inside the timed region)
        if pending is not None:
            merge_topk(*allgather_topk_finish(pending)); merges[0] += 1
        return users

    run_steps(0, args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.profile(True)
    launches0 = eng.launch_count
    merges[0] = 0
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    users_done = run_steps(args.warmup, args.steps)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    k_ms, k_n = eng.profile_read()
    eng.profile(False)
    launches = eng.launch_count - launches0 + merges[0]
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pairs = users_done * NI
    value = pairs / (ms * 1e-3)

    # ---- e2e: the public API with HOST buffers (user ids in, top-K lists out), copies inside the timed region
    h_users = [np.arange(((args.warmup + s) % n_blocks) * B, min(NU, (((args.warmup + s) % n_blocks) + 1) * B)) for s in range(args.steps)]
    def step_e2e(users_np):
        sc, ix = rec.recommend_all(users_np, top_k=TOP_K, filter_seen=True)
        if world > 1:
            all_s, all_i = allgather_topk(sc, ix)
            sc, ix = merge_topk(all_s, all_i)
        return sc.cpu(), ix.cpu()
    step_e2e(h_users[0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n_e2e = 0
    for u in h_users:
        hs, hi_ = step_e2e(u)
        n_e2e += len(u)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_val = n_e2e * NI / dt
    h2d = int(len(h_users[0]) * 8 + (len(h_users[0]) + 1) * 8)      # user indices + block CSR offsets (int64)
    d2h = int(len(h_users[0]) * TOP_K * 8)                          # fp32 score + int32 index per slot

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    H = list(spec.fusion_hidden_dims)
    wp = w_pair(fusion, spec.embedding_dim, H)
    pairs_per_launch = (users_done * (hi - lo)) / max(1, k_n)
    k_avg_ms = k_ms / max(1, k_n)
    achieved_tf = pairs_per_launch * wp / (k_avg_ms * 1e-3) / 1e12 if k_n else None
    peak_tf = pk["tf_sustained"]
    roofline = {"bound": "tensor", "kernel": f"pair-scoring ({eng.active_path})", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": (achieved_tf / peak_tf) if achieved_tf else None,
                "traffic": ncu_traffic(fusion, B // world, hi - lo),
                "peak_source": f"{pk['src']} bf16 sustained (kernel timed inside a long step); burst {pk['tf_burst']}",
                "flop_per_pair": wp, "pairs_per_launch": pairs_per_launch, "kernel_ms_avg": k_avg_ms, "kernel_launches": k_n,
                "kernel_share_of_step": (k_ms / ms) if ms else None}
    cpu = cpu_arm(args.config, args.cpu_seconds, args.seed) if (world == 1 and args.cpu_seconds > 0) else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if eng.active_path == "tcgen05" else "f32", "data": "synthetic",
            "users_per_sec": value / NI,
            "config": {"workload": desc, "n_users": NU, "n_items": NI, "fusion": fusion, "top_k": TOP_K,
                       "embedding_dim": spec.embedding_dim, "hidden": H, "users_per_step": B,
                       "items_per_rank": hi - lo, "parallelism": f"item-shard x{world}, all-gather of step s overlapped with scoring of step s+1" if world > 1 else "single GPU",
                       "kernel_path": eng.active_path, "filter_seen": True,
                       "l2": "flushed between steps by a 256 MiB memset inside the timed region"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "FastRecommender.recommend_all(host user ids) -> top-K lists copied to host"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _LazyIds:
    """Zero-padded id strings without materialising millions of them."""

    def __init__(self, n, prefix):
        self.n, self.prefix = n, prefix

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return f"{self.prefix}{int(i):08d}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="B", choices=list(CONFIGS))
    ap.add_argument("--user-block", type=int, default=4096, help="users per step per GPU")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample length (0 = skip)")
    ap.add_argument("--seed", type=int, default=20261018)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print(f"bench.py: note: warm-up {args.warmup} < 3 is below the timing-hygiene minimum", file=sys.stderr)
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()

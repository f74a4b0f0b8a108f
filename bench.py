#!/usr/bin/env python
"""bench.py — headline benchmark of the full-catalogue scoring + top-K path.

A "step" is one pass of the hot path over one block of users: every user of the
block scored against EVERY item of the catalogue, seen items filtered, top-50
kept (pxr_score_topk in its default exact mode: the fused 16-bit kernel keeps 64
candidates per user, they are re-scored in fp32 and re-ranked; with N>1 GPUs each
rank owns a contiguous item shard and the per-shard lists are exchanged with one
all-to-all over NVLink and merged by the rank that owns the user).

Workload at N=1 = BASELINE.json configs[1]: gated fusion, CLIP-512 + SBERT-384
cached features, Pixel200K-shaped synthetic (200 000 users x 96 282 items), top-50.
Per-GPU work is fixed as N grows (users per step = user_block * N, items per
rank = NI / N): "scaling": "weak".  The same JSON object carries an `also` array
with short runs of the other BASELINE configs (configs[2] attention at every N,
the concat shape and gated fusion at embedding_dim 128 at N=1, configs[3] Pixel8M-shaped at N=8, and at N>1 the
user-axis-sharded variant of the headline for comparison), each with its own
roofline and clocks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

`--impl reference` times the reference's CPU path on the host cores for the same
metric and config: the UNMODIFIED reference module, pip-installed from
/root/reference into baseline/_ref by `__graft_entry__.build()` (git-ignored, travels
to the GPU box; oracle/reference_arm.py) -- `kind: "reference"`; only when that
install is absent does it fall back to the PyTorch-eager oracle port (`kind: "port"`).
The oracle / reference are only ever the CPU arm here; the product path never
touches them.
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
    "B128": (200_000, 96_282, "gated", "gated fusion at embedding_dim 128 at the configs[1] shape (200K x 96K), top-50 (a configs[4] sweep point: the wide gated front end)"),
    "C": (1_001_822, 100_541, "attention", "configs[2] attention + numerical, Pixel1M-shaped 1M x 100K"),
    "D": (8_886_078, 407_082, "gated", "configs[3] Pixel8M-shaped 8.9M x 407K item-sharded"),
}
CONFIG_DIMS = {"B128": 128}   # config name -> embedding_dim when it is not 64 (scripts/sweep.py registers its sweep points here)
TOP_K = 50
METRIC = "scored user-item pairs/sec (full catalogue, top-50 per user)"
UNIT = "pairs/s"
EMBEDDING_DIM, HIDDEN = 64, [512, 256, 128]


def config_dict(cfg_name: str, n_gpus: int, user_block: int, shard: str = "items"):
    """The workload description both arms print (identical dicts for the same command line)."""
    NU, NI, fusion, desc = CONFIGS[cfg_name]
    par = "single GPU" if n_gpus == 1 else (
        f"item-shard x{n_gpus}: one all-to-all of the per-shard top-K lists per step, overlapped with the scoring of the next step, "
        f"merge by the owning rank" if shard == "items" else f"user-shard x{n_gpus}: replicas of the catalogue, no exchange")
    return {"workload": desc, "n_users": NU, "n_items": NI, "fusion": fusion, "top_k": TOP_K, "embedding_dim": CONFIG_DIMS.get(cfg_name, EMBEDDING_DIM),
            "hidden": HIDDEN, "filter_seen": True, "users_per_step": user_block * n_gpus,
            "parallelism": f"B200 arm: {par}; CPU arm: one process, all host threads",
            "l2": "B200 arm: flushed between steps by a 256 MiB memset inside the timed region; CPU arm: not applicable"}


def w_pair(fusion: str, D: int, H):
    """Algorithmic tensor FLOPs per scored pair: the MLP tail
    2*(sum H_{l-1} H_l + H_L), plus the per-pair first layer 2*D*H1 when the
    fused vector depends on the pair (gated / attention)."""
    tail = 2 * (sum(a * b for a, b in zip(H[:-1], H[1:])) + H[-1])
    if fusion == "gated" and D != 64:
        # F_GATEDW (csrc/score_tc.cu): layer 1 is a gate-weighted sum of per-user / per-item partials on CUDA cores,
        # so only the tail runs on the tensor pipe -- count what the kernel executes, not SURVEY's 2*D*H1 estimate
        return tail
    return tail + (0 if fusion == "concatenate" else 2 * D * H[0])


def ncu_traffic(fusion: str, users_per_launch: int, items_per_rank: int):
    """DRAM bytes (read + write) of ONE launch of the fused kernel from the committed `ncu --set full` capture of this
    same workload shape (profiles/ncu_traffic.json, written by scripts/ncu_summary.py traffic); None when no capture matches."""
    p = REPO / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return None
    best = None
    for e in json.loads(p.read_text()).get("captures", []):
        if e["fusion"] == fusion and e["users_per_launch"] == users_per_launch and e["items_per_rank"] == items_per_rank:
            best = e["dram_bytes_per_launch"]          # later entries (newer rounds) win
    return best


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
# CPU arm: the oracle port (batched reference forward + stable top-K; literal per-user loop)
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
                                         cfg=cs.spec_cfg(spec), NU=NU, NI=NI, fusion=fusion, spec=spec)
    return _CPU_WL[(cfg_name, seed)]


def reference_model(wl):
    """The UNMODIFIED reference module (oracle/reference_arm.py: the pip install under baseline/_ref) with this workload's
    weights, or None when the install is absent (then the oracle port is timed and `kind` says "port")."""
    if "ref_model" not in wl:
        from oracle import reference_arm as ra
        wl["ref_model"] = ra.build_model(wl["spec"], wl["sd"]) if ra.available() else None
    return wl["ref_model"]


def cpu_arm(cfg_name: str, seconds: float, seed: int, first_user: int = 0):
    """Times the reference's CPU path on a bounded sample of the SAME workload: `n` users x the full catalogue as
    flattened pairs through `MultimodalRecommender.forward` (multimodal.py:528-610), seen items dropped, stable top-50
    (recommender.py:88-106).  kind "reference": the unmodified reference module installed under baseline/_ref
    (oracle/reference_arm.py); kind "port" (only when that install is absent): oracle/pxr_oracle_torch.py, the same forward
    restated op for op in PyTorch eager fp32 and pinned to the reference's outputs in tests/.  Returns pairs/s."""
    import torch
    from oracle import pxr_oracle_torch as ot
    from oracle import reference_arm as ra
    wl = cpu_workload(cfg_name, seed)
    NU, NI, fusion = wl["NU"], wl["NI"], wl["fusion"]
    threads = torch.get_num_threads()
    ref = reference_model(wl)

    def run(users):
        if ref is not None:
            ra.recommend_block(ref, torch.as_tensor(users), wl["feats"], TOP_K, wl["indptr"], wl["idx"])
        else:
            ot.recommend_block(wl["sd"], wl["cfg"], torch.as_tensor(users), wl["feats"], TOP_K, wl["indptr"], wl["idx"])

    t0 = time.perf_counter(); run(np.arange(first_user, first_user + 2) % NU); t1 = (time.perf_counter() - t0) / 2
    n = int(max(2, min(NU, seconds / max(t1, 1e-3))))
    users = np.arange(first_user + 2, first_user + 2 + n) % NU
    t0 = time.perf_counter(); run(users); dt = time.perf_counter() - t0
    what = ("the unmodified reference MultimodalRecommender.forward (pip-installed under baseline/_ref, cached features through "
            "identity backbones)" if ref is not None else "PyTorch-eager fp32 restatement of the reference forward")
    return dict(value=n * NI / dt, unit=UNIT, cores=threads, kind="reference" if ref is not None else "port", seconds=dt, users=n,
                sample=f"{n} users x {NI} items of config {cfg_name} ({fusion}): {what} on flattened pairs "
                       f"({threads} intra-op threads of {os.cpu_count()} cores) + seen filter + stable top-{TOP_K}")


def cpu_literal(cfg_name: str, seconds: float, seed: int, max_items: int = 20_000):
    """The LITERAL path a reference user runs: one `Recommender.get_recommendations(user,
    top_k=50, filter_seen=True)` call per user -- per-call id lists, 256-item batches, per-item feature-dict fetches,
    torch.stack, sklearn LabelEncoder.transform, one forward per batch, Python sort (oracle/pxr_oracle_torch.py::
    LiteralRecommender, src/inference/recommender.py:52-236 step by step; pinned to the real reference's lists in tests/).
    The per-item feature dicts of a big catalogue take minutes to build, so the catalogue is cut to its first
    `max_items` items (the cost per pair does not depend on the catalogue size beyond that)."""
    import torch
    from sklearn.preprocessing import LabelEncoder
    from oracle import pxr_oracle_torch as ot
    from pixelrec_multimodal_b200 import synthetic as syn
    wl = cpu_workload(cfg_name, seed)
    NU, NI = wl["NU"], min(wl["NI"], max_items)
    n_enc_users = min(NU, 20_000)                         # the per-call user id list is O(n_users) (recommender.py:64)
    feats, indptr, idx = wl["feats"], wl["indptr"].numpy(), wl["idx"].numpy()
    iids, uids = syn.item_ids(NI), syn.user_ids(n_enc_users)

    class _DS:
        pass
    ds = _DS()
    ds.user_encoder, ds.item_encoder = LabelEncoder().fit(uids), LabelEncoder().fit(iids)
    ds.feature_cache = {}
    ones = torch.ones(1, dtype=torch.long)
    for i, iid in enumerate(iids):
        d = {"tag_idx": feats["tag_idx"][i]}
        if "vis" in feats:
            d["image"] = feats["vis"][i]
        if "txt" in feats:
            d["text_input_ids"] = feats["txt"][i]
            d["text_attention_mask"] = ones
        if "num" in feats:
            d["numerical_features"] = feats["num"][i]
        ds.feature_cache[iid] = d
    ds.get_user_history = lambda uid: {iids[int(j)] for j in idx[indptr[int(uid[1:])]:indptr[int(uid[1:]) + 1]] if j < NI}
    sd = dict(wl["sd"])
    sd["item_embedding.weight"] = sd["item_embedding.weight"][:NI]
    ref = reference_model(wl)
    if ref is not None:          # the installed reference Recommender on the reference model cut to the same catalogue
        from oracle import reference_arm as ra
        import dataclasses
        sub = ra.build_model(dataclasses.replace(wl["spec"], n_items=NI), sd)
        lit = ra.literal_recommender(sub, ds)
    else:
        lit = ot.LiteralRecommender(sd, wl["cfg"], ds)
    t0 = time.perf_counter(); lit.get_recommendations(uids[0], top_k=TOP_K, filter_seen=True); t1 = time.perf_counter() - t0
    n = int(max(1, min(64, seconds / max(t1, 1e-3))))
    t0 = time.perf_counter()
    for u in range(1, 1 + n):
        r = lit.get_recommendations(uids[u], top_k=TOP_K, filter_seen=True)
        assert len(r) == TOP_K
    dt = time.perf_counter() - t0
    return dict(value=n * NI / dt, users_per_sec=n / dt, users=n, items=NI, seconds=dt, kind="reference" if ref is not None else "port",
                sample=f"{n} {'reference Recommender' if ref is not None else 'LiteralRecommender (port)'}.get_recommendations calls "
                       f"(top-{TOP_K}, filter_seen) over the first {NI} items of config {cfg_name}, {n_enc_users} users in the encoder")


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
    NI = CONFIGS[args.config][1]
    base = {"value": v, "unit": UNIT, "cores": vals[-1]["cores"], "kind": vals[-1]["kind"], "sample": vals[-1]["sample"]}
    if args.literal_seconds > 0:
        lit = cpu_literal(args.config, args.literal_seconds, args.seed)
        base.update(literal_value=lit["value"], literal_users_per_sec=lit["users_per_sec"], literal_sample=lit["sample"])
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, len(vals)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "users_per_sec": v / NI,
            "config": config_dict(args.config, args.gpus, args.user_block),
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
class _LazyIds:
    """Zero-padded id strings without materialising millions of them."""

    def __init__(self, n, prefix):
        self.n, self.prefix = n, prefix

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if not 0 <= int(i) < self.n:
            raise IndexError(i)
        return f"{self.prefix}{int(i):08d}"


def run_config(args, cfg_name: str, steps: int, warmup: int, world: int, rank: int, dev, shard: str = "items",
               with_e2e: bool = True, with_cpu: bool = False, with_checks: bool = True):
    """One measured run of `cfg_name` on the current process group; returns the JSON line as a dict (every rank)."""
    import torch
    import torch.distributed as dist
    from pixelrec_multimodal_b200 import FastMultimodalRecommender, FastRecommender, ItemFeatureStore, synthetic as syn
    from pixelrec_multimodal_b200.engine import merge_topk
    from pixelrec_multimodal_b200.sharding import exchange_owned_finish, exchange_owned_start, owned_slice, shard_range

    NU, NI, fusion, desc = CONFIGS[cfg_name]
    spec = syn.ModelSpec(n_users=NU, n_items=NI, fusion_type=fusion, embedding_dim=CONFIG_DIMS.get(cfg_name, EMBEDDING_DIM))
    sd, feats, hist = syn.torch_workload(spec, dev, seed=args.seed)      # identical on every rank (same seed)
    # BatchNorm statistics matched to the activations and logits spread to std 2, like a trained checkpoint
    syn.condition_like_trained(sd, spec, feats)

    def make_model():
        m = FastMultimodalRecommender(
            n_users=NU, n_items=NI, n_tags=spec.n_tags, num_numerical_features=spec.num_numerical_features,
            embedding_dim=spec.embedding_dim, vision_model_name=f"cached{spec.vision_dim}",
            language_model_name=f"cached{spec.language_dim}", use_contrastive=False, fusion_type=fusion,
            fusion_hidden_dims=list(spec.fusion_hidden_dims), kernel_path=args.path).to(dev)
        m.load_state_dict(sd, strict=False)
        return m

    class _Enc:
        def __init__(self, n, p): self.classes_ = _LazyIds(n, p)

    class _DS:
        user_encoder, item_encoder, interactions = _Enc(NU, "u"), _Enc(NI, "i"), None

    model = make_model()
    item_sharded = world > 1 and shard == "items"
    lo, hi = shard_range(NI, world, rank) if item_sharded else (0, NI)
    store = ItemFeatureStore(feats["tag_idx"], feats["vis"], feats["txt"], feats["num"])
    B = args.user_block * world                                   # users per step (whole job)
    rec = FastRecommender(model, _DS(), dev, item_features=store, n_users=NU, n_items=NI,
                          history=(hist["train_indptr"], hist["train_idx"]), item_range=(lo, hi), user_block=B)
    eng = rec.engine()
    # exact mode across item shards: the shards exchange their raw 64-slot lists and the rank that owns a user re-scores the
    # merged candidates against fp32 records of the whole catalogue (1 280 B per item, kept by every rank)
    rs_eng = rec.rescore_engine() if item_sharded else None
    K_EX = 64 if item_sharded else TOP_K
    if item_sharded:
        eng.set_rescore(False)
    torch.cuda.synchronize()

    n_blocks = max(1, NU // B)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    uemb = model.user_embedding.weight.detach()
    d_indptr, d_idx = rec.device_history()
    merges = [0]
    all_users = torch.arange(NU, device=dev)

    def block_users(s):
        """(first user, count) this rank scores in step s: the whole block when items are sharded, its 1/world slice
        of the block under user-axis sharding"""
        b = s % n_blocks
        u0 = b * B
        n = min(NU, u0 + B) - u0
        if world > 1 and not item_sharded:
            a, z = owned_slice(n, world, rank)
            return u0 + a, z - a, n
        return u0, n, n

    def score_block(s):
        """inputs already in HBM: the block's user indices and the resident history CSR"""
        u0, n, n_job = block_users(s)
        sc, ix = eng.score_topk(uemb, all_users[u0:u0 + n], K_EX, d_indptr[u0:u0 + n + 1], d_idx)
        return sc, ix, n_job

    def finish_owned(pending):
        """merge the shards' raw lists of the users this rank owns, then the fp32 re-score of the merged candidates"""
        handle, u0, n = pending
        ms_, mi_ = merge_topk(*exchange_owned_finish(handle))
        a, z = owned_slice(n, world, rank)
        merges[0] += 1
        return rs_eng.rescore_topk(uemb, all_users[u0 + a:u0 + z], mi_, TOP_K) if z > a else (ms_[:, :TOP_K], mi_[:, :TOP_K])

    def run_steps(first, count):
        """`count` steps.  Item shards: the all-to-all of step s (one packed collective on NCCL's stream) overlaps the
        scoring kernel of step s + 1; the owning rank's merge is enqueued behind that kernel."""
        users = 0
        pending = None
        for s in range(first, first + count):
            sc, ix, n = score_block(s)
            users += n
            if item_sharded:
                handle = exchange_owned_start(sc, ix)
                if pending is not None:
                    finish_owned(pending)
                pending = (handle, block_users(s)[0], n)
            flush.zero_()                                         # L2 flush between steps (inside the timed region)
        if pending is not None:
            finish_owned(pending)
        return users

    run_steps(0, warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.profile(True)
    launches0 = eng.launch_count + (rs_eng.launch_count if rs_eng is not None else 0)
    merges[0] = 0
    sampler = ClockSampler(dev.index)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    users_done = run_steps(warmup, steps)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    k_ms, k_n = eng.profile_read()
    eng.profile(False)
    launches = eng.launch_count + (rs_eng.launch_count if rs_eng is not None else 0) - launches0 + merges[0]
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pairs = users_done * NI
    value = pairs / (ms * 1e-3)

    # ---- e2e: the public API with HOST buffers (user ids in, top-K lists out), copies inside the timed region
    e2e = None
    if with_e2e:
        def host_users(s):
            u0, n, _ = block_users(warmup + s)
            return np.arange(u0, u0 + n)
        h_users = [host_users(s) for s in range(steps)]

        def step_e2e(users_np):
            sc, ix = rec.recommend_all(users_np, top_k=K_EX, filter_seen=True)
            if item_sharded:                                      # every rank scored the whole block against its shard
                sc, ix = finish_owned((exchange_owned_start(sc, ix), int(users_np[0]), len(users_np)))
            return sc.cpu(), ix.cpu()                             # the lists of the users this rank owns
        step_e2e(h_users[0])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_e2e = 0
        for s, u in enumerate(h_users):
            step_e2e(u)
            n_e2e += block_users(warmup + s)[2]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        n_job = block_users(warmup)[2]
        e2e = {"value": n_e2e * NI / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(world * (len(h_users[0]) * 8 + (len(h_users[0]) + 1) * 8)),   # user indices + block CSR offsets (int64), every rank
               "d2h_bytes_per_step": int(n_job * TOP_K * 8),                                           # fp32 score + int32 index per slot
               "api": "FastRecommender.recommend_all(host user ids) -> top-K lists copied to host" +
                      (" (each rank: the lists of the users it owns after the all-to-all)" if item_sharded else "")}

    # ---- checks outside the timed region
    checks = {}
    if with_checks and eng.active_path == "tcgen05" and not item_sharded:
        # raw 16-bit lists vs the exact-mode lists (fp32 re-scored) of the same users on this rank's item range
        u0, n, _ = block_users(warmup)
        nu = min(256, n)
        uu = all_users[u0:u0 + nu]
        xs, xi = eng.score_topk(uemb, uu, TOP_K, d_indptr[u0:u0 + nu + 1], d_idx)
        eng.set_rescore(False)
        rs, ri = eng.score_topk(uemb, uu, TOP_K, d_indptr[u0:u0 + nu + 1], d_idx)
        eng.set_rescore(True)
        ov = (xi.unsqueeze(2) == ri.unsqueeze(1)).any(dim=2).sum(dim=1).float()
        fp32_of_raw = eng.score_pairs(uemb, uu.repeat_interleave(TOP_K), (ri.reshape(-1).clamp_min(lo) - lo).to(torch.int64)).view(nu, TOP_K)
        checks["parity_sample"] = {
            "users": int(nu), "items": int(hi - lo),
            "raw16_vs_exact_top50_overlap_mean": float(ov.mean()), "raw16_vs_exact_top50_overlap_min": float(ov.min()),
            "raw16_max_abs_score_error_vs_fp32": float((rs - fp32_of_raw)[ri >= 0].abs().max()),
            "note": "exact mode (default): the 64 candidates the 16-bit kernel keeps per user are re-scored with the fp32 arithmetic of "
                    "forward() and re-ranked; tests/test_gpu_parity.py::test_catalogue_scale_parity checks both against the exact forward"}
    if with_checks and item_sharded:
        # merged sharded lists == the unsharded lists of rank 0, bit for bit: the raw 16-bit lists (the per-pair arithmetic does
        # not depend on the tiling) and the exact-mode lists (same global candidates, same fp32 records)
        from pixelrec_multimodal_b200.sharding import gather_owned
        nu = 64
        uu = all_users[:nu]
        sc, ix = eng.score_topk(uemb, uu, K_EX, d_indptr[:nu + 1], d_idx)
        ms_, mi_ = merge_topk(*exchange_owned_finish(exchange_owned_start(sc, ix)))
        a, z = owned_slice(nu, world, rank)
        xs_, xi_ = rs_eng.rescore_topk(uemb, uu[a:z], mi_, TOP_K)
        raw_s, raw_i = gather_owned(ms_, mi_, nu)
        ex_s, ex_i = gather_owned(xs_, xi_, nu)
        if rank == 0:
            full_model = make_model()
            full = FastRecommender(full_model, _DS(), dev, item_features=store, n_users=NU, n_items=NI,
                                   history=(hist["train_indptr"], hist["train_idx"]), user_block=B)
            fe = full.engine()
            fu = full_model.user_embedding.weight.detach()
            fe.set_rescore(False)
            fs, fi = fe.score_topk(fu, uu, K_EX, d_indptr[:nu + 1], d_idx)
            raw_ok = bool(torch.equal(fi, raw_i) and torch.equal(fs, raw_s))
            fe.set_rescore(True)
            fs, fi = fe.score_topk(fu, uu, TOP_K, d_indptr[:nu + 1], d_idx)
            exact_ok = bool(torch.equal(fi, ex_i) and torch.equal(fs, ex_s))
            checks["shard_check"] = "ok" if (raw_ok and exact_ok) else "MISMATCH"
            checks["shard_check_detail"] = {"users": nu, "raw16_top64_lists_bit_identical": raw_ok, "exact_top50_lists_bit_identical": exact_ok,
                                            "exact_users_identical": int((fi == ex_i).all(dim=1).sum())}
            del full, fe, full_model
        dist.barrier()

    pk = peaks()
    H = list(spec.fusion_hidden_dims)
    wp = w_pair(fusion, spec.embedding_dim, H)
    rank_users = users_done if (world == 1 or item_sharded) else users_done / world
    pairs_per_launch = (rank_users * (hi - lo)) / max(1, k_n)
    k_avg_ms = k_ms / max(1, k_n)
    achieved_tf = pairs_per_launch * wp / (k_avg_ms * 1e-3) / 1e12 if k_n else None
    peak_tf = pk["tf_sustained"]
    roofline = {"bound": "tensor", "kernel": f"pair-scoring ({eng.active_path})", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": (achieved_tf / peak_tf) if achieved_tf else None,
                "traffic": ncu_traffic(fusion if (spec.embedding_dim == 64 or fusion == "concatenate") else f"{fusion}_d{spec.embedding_dim}",
                                       int(rank_users / max(1, steps)), hi - lo),    # captures are per kernel: gated at D != 64 is another front end
                "peak_source": f"{pk['src']} bf16 sustained (kernel timed inside a long step); burst {pk['tf_burst']}",
                "flop_per_pair": wp, "pairs_per_launch": pairs_per_launch, "kernel_ms_avg": k_avg_ms, "kernel_launches": k_n,
                "kernel_share_of_step": (k_ms / ms) if ms else None}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if eng.active_path == "tcgen05" else "f32", "data": "synthetic",
            "users_per_sec": value / NI,
            "config": config_dict(cfg_name, world, args.user_block, shard),
            "run": {"items_per_rank": hi - lo, "kernel_path": eng.active_path, "exact_rescore": bool(eng.rescore) or item_sharded,
                    "shard_axis": shard if world > 1 else None},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if e2e is not None:
        line["e2e"] = e2e
    line.update(checks)
    if with_cpu and rank == 0:
        cpu = cpu_arm(cfg_name, args.cpu_seconds, args.seed)
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if args.literal_seconds > 0:
            lit = cpu_literal(cfg_name, args.literal_seconds, args.seed)
            line["cpu_baseline"].update(literal_value=lit["value"], literal_users_per_sec=lit["users_per_sec"], literal_sample=lit["sample"])
    if with_cpu and rank == 0 and world == 1:
        # the reference's own calling pattern on this arm: one get_recommendations(user_id: str) call per user
        # (src/inference/recommender.py:52-110), host strings in, list of (item id, score) out
        n_calls = 200
        ids = [f"u{u:08d}" for u in range(n_calls)]
        rec.get_recommendations(ids[0], top_k=TOP_K, filter_seen=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = 0
        for uid in ids:
            got += len(rec.get_recommendations(uid, top_k=TOP_K, filter_seen=True))
        dt = time.perf_counter() - t0
        line["literal_api"] = {"api": "FastRecommender.get_recommendations(user_id: str, top_k=50, filter_seen=True), one call per user",
                               "calls": n_calls, "lists_returned": got // TOP_K, "ms_per_call": dt / n_calls * 1e3,
                               "users_per_sec": n_calls / dt, "value": n_calls * NI / dt, "unit": UNIT}
    # free this configuration's device memory before the next one
    del rec, eng, rs_eng, model, store, feats, hist, sd, flush, uemb, d_indptr, d_idx, all_users
    torch.cuda.empty_cache()
    return line


def b200_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = run_config(args, args.config, args.steps, args.warmup, world, rank, dev, shard=args.shard, with_e2e=True,
                      with_cpu=(world == 1 and args.cpu_seconds > 0))
    also = []
    if args.also and args.config == "B":
        extra = [("C", "items")] + ([("Bc", "items"), ("B128", "items")] if world == 1 else [("B", "users")]) + ([("D", "items")] if world == 8 else [])
        for cfg_name, shard in extra:
            try:
                r = run_config(args, cfg_name, args.also_steps, 3, world, rank, dev, shard=shard, with_e2e=False, with_checks=(shard == "items"))
                also.append({k: r[k] for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "users_per_sec", "dtype", "config",
                                               "run", "gpu_launches", "clocks", "roofline") + tuple(c for c in ("parity_sample", "shard_check", "shard_check_detail") if c in r)})
            except Exception as e:                                  # an extra run must never take the headline line down
                also.append({"config": config_dict(cfg_name, world, args.user_block, shard), "error": f"{type(e).__name__}: {e}"[:300]})
                torch.cuda.empty_cache()
    if rank == 0:
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="B", choices=list(CONFIGS))
    ap.add_argument("--user-block", type=int, default=4096, help="users per step per GPU")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--shard", default="items", choices=["items", "users"],
                    help="N > 1: item-axis shards + exchange (north_star) or user-axis shards (replicas, no exchange)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample length (0 = skip)")
    ap.add_argument("--literal-seconds", type=float, default=8.0, help="literal per-user CPU baseline sample length (0 = skip)")
    ap.add_argument("--no-also", dest="also", action="store_false", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--also-steps", type=int, default=5)
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

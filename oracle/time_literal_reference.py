"""Cross-check of bench.py's literal CPU baseline: time the UNMODIFIED reference `Recommender.get_recommendations`
(src/inference/recommender.py:52-110, imported from /root/reference with the stub backbones of oracle/make_golden.py) next to
oracle/pxr_oracle_torch.py::LiteralRecommender (the port that travels to the GPU box) on the same light dataset, same users,
same machine, and check that they return the same lists.  TEST INFRASTRUCTURE ONLY; needs /root/reference (build container).

  python oracle/time_literal_reference.py > profiles/r02_literal_reference_check.json
"""
from __future__ import annotations

import json
import logging
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from tests import _cases as cs                  # noqa: E402  (this repo's tests package, before the reference's shadows it)
from oracle import make_golden as mg            # noqa: E402  (puts /root/reference on sys.path)
from oracle import pxr_oracle_torch as ot       # noqa: E402
from pixelrec_multimodal_b200 import synthetic as syn  # noqa: E402


def main():
    from src.inference.recommender import Recommender
    logging.disable(logging.CRITICAL)
    out = {"host_cores": os.cpu_count(), "torch_threads": torch.get_num_threads(), "cases": []}
    for fusion, n_items in (("concatenate", 2000), ("gated", 2000), ("attention", 2000)):
        spec = syn.ModelSpec(n_users=200, n_items=n_items, fusion_type=fusion)       # configs[0] catalogue size
        sd, feats = cs.make_workload(spec, syn.SEED + 71)
        indptr, idx, _ = syn.make_histories(spec.n_users, spec.n_items, seed=7, lo=3, hi=30)
        ds = cs.LightDataset(spec, feats, indptr, idx)
        model = mg.build_reference_model(spec, sd, double=False)
        ref = Recommender(model, ds, torch.device("cpu"))
        ref._debug_has_run_recommender = True                                        # skip the debug forward (recommender.py:193-219)
        lit = ot.LiteralRecommender({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, cs.spec_cfg(spec), ds)
        users = ds.uids[:16]
        ref.get_recommendations(users[0], top_k=50); lit.get_recommendations(users[0], top_k=50)      # warm-up
        t0 = time.perf_counter(); r_lists = [ref.get_recommendations(u, top_k=50, filter_seen=True) for u in users]; t_ref = time.perf_counter() - t0
        t0 = time.perf_counter(); p_lists = [lit.get_recommendations(u, top_k=50, filter_seen=True) for u in users]; t_port = time.perf_counter() - t0
        same = all([a[0] for a in x] == [b[0] for b in y] for x, y in zip(r_lists, p_lists))
        dmax = max(abs(a[1] - b[1]) for x, y in zip(r_lists, p_lists) for a, b in zip(x, y))
        out["cases"].append({"fusion": fusion, "n_items": n_items, "users": len(users),
                             "reference_users_per_s": len(users) / t_ref, "port_users_per_s": len(users) / t_port,
                             "reference_pairs_per_s": len(users) * n_items / t_ref, "port_pairs_per_s": len(users) * n_items / t_port,
                             "lists_identical": bool(same), "max_abs_score_diff": float(dmax)})
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()

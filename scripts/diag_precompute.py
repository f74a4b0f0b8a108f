#!/usr/bin/env python
"""Per-call device (CUDA events) and host (perf_counter) time of precompute_items, 10 calls in a row."""
import sys, time
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from pixelrec_multimodal_b200 import FastMultimodalRecommender, synthetic as syn

fusion = sys.argv[1] if len(sys.argv) > 1 else "concatenate"
dev = torch.device("cuda:0")
spec = syn.ModelSpec(n_users=4096, n_items=96282, fusion_type=fusion)
sd, feats, _ = syn.torch_workload(spec, dev, seed=1, with_histories=False)
m = FastMultimodalRecommender(n_users=spec.n_users, n_items=spec.n_items, n_tags=spec.n_tags, num_numerical_features=7,
                              embedding_dim=64, vision_model_name="cached512", language_model_name="cached384", fusion_type=fusion).to(dev)
m.load_state_dict(sd, strict=False)
e = m.engine("catalogue")
w = m.item_embedding.weight.detach()
for i in range(10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    e.precompute_items(w, feats["tag_idx"], feats["vis"], feats["txt"], feats["num"])
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{fusion} call {i}: device {e0.elapsed_time(e1):.3f} ms, host enqueue {1e3 * (t1 - t0):.3f} ms, host to sync {1e3 * (t2 - t0):.3f} ms, "
          f"reserved {torch.cuda.memory_reserved() >> 20} MiB", flush=True)

"""Debug / accuracy probe of the tcgen05 scoring path against the CPU oracle (GPU box only)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import pxr_oracle as orc
from pixelrec_multimodal_b200 import synthetic as syn
from tests import _cases as cs


def run(n_users, n_items, k, filt, fusion="gated", seed=0, verbose=True):
    spec = syn.ModelSpec(n_users=n_users, n_items=n_items, fusion_type=fusion)
    sd, feats = cs.make_workload(spec, syn.SEED + seed)
    indptr, idx, _ = syn.make_histories(n_users, n_items, seed=syn.SEED + seed, lo=3, hi=min(40, max(4, n_items // 3)))
    model = cs.torch_model_from(spec, sd, kernel_path="tcgen05")
    eng = model.engine("catalogue")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    eng.precompute_items(model.item_embedding.weight.detach(), t(feats["tag_idx"]), t(feats["vis"]), t(feats["txt"]), t(feats["num"]))
    assert eng.active_path == "tcgen05", eng.active_path
    users = torch.arange(n_users).cuda()
    args = (t(indptr), t(idx)) if filt else ()
    t0 = time.time()
    s, i = eng.score_topk(model.user_embedding.weight.detach(), users, k, *args)
    torch.cuda.synchronize()
    dt = time.time() - t0
    s, i = s.cpu().numpy().astype(np.float64), i.cpu().numpy()
    ref = orc.score_block(sd, cs.spec_cfg(spec), np.arange(n_users), 0, n_items, feats)
    refz = orc.score_block(sd, cs.spec_cfg(spec), np.arange(n_users), 0, n_items, feats, return_logit=True)
    max_err, max_zerr, mism, total, bad = 0.0, 0.0, 0, 0, 0
    for u in range(n_users):
        seen = idx[indptr[u]:indptr[u + 1]] if filt else None
        rsel, rsc = orc.topk_from_scores(ref[u], k, seen=seen)
        n = len(rsel)
        gi, gs = i[u][:n], s[u][:n]
        if not (np.all(i[u][n:] == -1) and np.all(gi >= 0)):
            bad += 1
            if verbose and bad < 4:
                print("user", u, "bad padding / count: got", int((i[u] >= 0).sum()), "want", n, i[u][:8], rsel[:8])
            continue
        err = np.abs(gs - ref[u][gi])
        max_err = max(max_err, float(err.max()))
        with np.errstate(divide="ignore"):
            gz = np.log(np.clip(gs, 1e-12, 1 - 1e-12) / (1 - np.clip(gs, 1e-12, 1 - 1e-12)))
        zsel = np.abs(refz[u][gi]) < 8
        if zsel.any():
            max_zerr = max(max_zerr, float(np.abs(gz - refz[u][gi])[zsel].max()))
        mism += int(np.sum(gi != rsel)); total += n
        if seen is not None and set(gi.tolist()) & set(seen.tolist()):
            bad += 1; print("user", u, "returned a seen item")
    print(f"[{fusion} NU={n_users} NI={n_items} K={k} filter={filt}] time {dt*1e3:.1f} ms  max|ds|={max_err:.3e}  "
          f"max|dz|={max_zerr:.3e}  positions differing {mism}/{total}  bad users {bad}")
    return max_err, mism, total, bad


if __name__ == "__main__":
    run(16, 48, 64, False)
    run(16, 48, 64, True)
    run(40, 1000, 50, True, seed=1)
    run(300, 5003, 50, True, seed=2)

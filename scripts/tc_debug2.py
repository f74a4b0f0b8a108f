import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import pxr_oracle as orc
from pixelrec_multimodal_b200 import synthetic as syn
from tests import _cases as cs
from tests.test_gpu_parity import _tc_workload, _engine_for, _lowp_scores

n_users, n_items, k = int(sys.argv[1]), int(sys.argv[2]), 50
spec, sd, feats, indptr, idx, _ = _tc_workload(n_users, n_items, syn.SEED + 22)
model, eng = _engine_for(spec, sd, feats, "tcgen05")
users = torch.arange(n_users).cuda()
fs, fi = eng.score_topk(model.user_embedding.weight.detach(), users, k, torch.from_numpy(indptr).cuda(), torch.from_numpy(idx).cuda())
torch.cuda.synchronize()
s, i = fs.cpu().numpy().astype(np.float64), fi.cpu().numpy()
sample = np.unique(np.concatenate([np.arange(0, 40), np.arange(1176, 1200), np.arange(2040, 2056), np.arange(n_users - 24, n_users)]))
sample = sample[sample < n_users]
emu = _lowp_scores(sd, spec, feats, sample)
for r, u in enumerate(sample):
    seen = idx[indptr[u]:indptr[u + 1]]
    rsel, rsc = orc.topk_from_scores(emu[r], k, seen=seen)
    gi, gs = i[u], s[u]
    ok = gi >= 0
    err = np.abs(gs[ok] - emu[r][gi[ok]])
    nbad = int((err > 5e-4).sum())
    miss = set(rsel.tolist()) - set(gi.tolist())
    if nbad or miss or (~ok).any():
        print(f"user {u} (group {u//16} cta {(u//8)%2} slot {u%8}): bad scores {nbad} max err {err.max():.3e} missing {sorted(miss)[:8]} "
              f"npad {int((~ok).sum())} bad idx {gi[ok][err > 5e-4][:8]} got {gs[ok][err>5e-4][:4]} want {emu[r][gi[ok]][err>5e-4][:4]}")
print("done; checked", len(sample))

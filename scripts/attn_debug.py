import numpy as np, torch, sys
sys.path.insert(0, '.')
from oracle import pxr_oracle as orc
from pixelrec_multimodal_b200 import synthetic as syn
from tests import _cases as cs
from tests.test_gpu_parity import _tc_workload, _engine_for, _lowp_scores
for (n_users, n_items, k) in [(16, 48, 64), (48, 1500, 50)]:
    spec, sd, feats, indptr, idx, _ = _tc_workload(n_users, n_items, syn.SEED + 21, "attention")
    model, eng = _engine_for(spec, sd, feats, "tcgen05", dtype="bf16")
    users = np.arange(n_users)
    s, i = eng.score_topk(model.user_embedding.weight.detach(), torch.from_numpy(users).cuda(), min(k, 64))
    s, i = s.cpu().numpy().astype(np.float64), i.cpu().numpy()
    emu = _lowp_scores(sd, spec, feats, users, orc.round_bf16)
    ref = orc.score_block(sd, cs.spec_cfg(spec), users, 0, n_items, feats)
    errs = []
    for u in users:
        v = i[u] >= 0
        e = np.abs(s[u][v] - emu[u][i[u][v]])
        errs.append(e.max())
        if e.max() > 1e-3:
            j = np.argmax(e); print("user", u, "item", i[u][v][j], "got", s[u][v][j], "emu", emu[u][i[u][v][j]], "exact", ref[u][i[u][v][j]])
    print(n_users, n_items, "max err vs emu", max(errs), "emu-vs-exact", np.abs(emu - ref).max())

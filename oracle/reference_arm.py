"""The UNMODIFIED reference as bench.py's CPU arm.  TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under
pixelrec_multimodal_b200/ imports this module; bench.py's `--impl reference` leg and its `cpu_baseline` field are the only callers.

The reference (Joacodef/PixelRec_Multimodal) is pure Python / PyTorch.  `install()` is the committed recipe that makes it
travel to the GPU box: the contract's one offline install,

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <a copy of /root/reference under /tmp>

(`--no-deps`: matplotlib, a requirement of setup.py the hot path never imports, is not in the wheelhouse; the copy under /tmp
because /root/reference is read-only and setuptools writes build/ and *.egg-info into the source tree).  `baseline/_ref` is
git-ignored -- no reference source enters the history -- but not gpurun-ignored, so the installed package is on the box.
`__graft_entry__.build()` runs it when /root/reference is present.

With the install present, `kind` of the CPU arm is "reference":
  * batched: `MultimodalRecommender.forward` (src/models/multimodal.py:528-610) of the installed package on flattened
    (user, item) pairs of the SAME workload -- cached CLIP-512 / SBERT-384 features fed through identity backbones registered
    in the reference's own MODEL_CONFIGS registry (the recipe the goldens were generated with) --
    seen items dropped, stable descending sort, first 50 (src/inference/recommender.py:88-106);
  * literal: the installed `Recommender.get_recommendations(user, top_k=50, filter_seen=True)`, one call per user
    (src/inference/recommender.py:52-110).
Without it (never the case after build()), bench.py falls back to the oracle port (`kind` "port").
"""
from __future__ import annotations

import logging
import os
import shutil
import subprocess
import sys
import tempfile
import types
from pathlib import Path
from typing import Dict, Optional

REPO = Path(__file__).resolve().parent.parent
REF_DIR = REPO / "baseline" / "_ref"
REF_SRC = Path(os.environ.get("PXR_REFERENCE", "/root/reference"))


def available() -> bool:
    return (REF_DIR / "src" / "models" / "multimodal.py").exists() and (REF_DIR / "src" / "inference" / "recommender.py").exists()


def install(force: bool = False) -> str:
    """The install recipe above; returns a one-line outcome (recorded in DESIGN.md section 7)."""
    if available() and not force:
        return "present"
    if not (REF_SRC / "setup.py").exists():
        return f"skipped: {REF_SRC} not present (GPU box: the package installed in the build container travels in baseline/_ref)"
    tmp = Path(tempfile.mkdtemp(prefix="pxr_ref_src_"))
    try:
        src = tmp / "reference"
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        if REF_DIR.exists():
            shutil.rmtree(REF_DIR)
        REF_DIR.parent.mkdir(parents=True, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", str(REF_DIR), str(src)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=str(tmp))
        if r.returncode != 0 or not available():
            return "failed: " + " | ".join(r.stdout.strip().splitlines()[-3:])
        return "installed"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _import_reference():
    """Puts baseline/_ref BEHIND the repo on sys.path (the wheel also ships a top-level `tests` package that must not shadow
    this repo's) and returns (multimodal module, MODEL_CONFIGS, Recommender class)."""
    p = str(REF_DIR)
    if p not in sys.path:
        sys.path.append(p)
    import src.models.multimodal as mm
    from src.config import MODEL_CONFIGS
    from src.inference.recommender import Recommender
    if not str(Path(mm.__file__).resolve()).startswith(str(REF_DIR.resolve())) and not str(Path(mm.__file__).resolve()).startswith(str(REF_SRC)):
        raise RuntimeError(f"`src` resolved to {mm.__file__}, not to the reference install")
    return mm, MODEL_CONFIGS, Recommender


class _Out:
    def __init__(self, x):
        self.pooler_output = x
        self.last_hidden_state = None


def build_model(spec, sd: Dict[str, "object"]):
    """The reference `MultimodalRecommender` with this workload's weights (strict load) and cached-feature backbones."""
    import numpy as np
    import torch
    mm, MODEL_CONFIGS, _ = _import_reference()

    class _VisionStub(torch.nn.Module):
        def forward(self, pixel_values=None, **kw):
            return _Out(pixel_values)

    class _LanguageStub(torch.nn.Module):
        def forward(self, input_ids=None, attention_mask=None, **kw):
            return _Out(input_ids)

    vkey, lkey = f"cached{spec.vision_dim}", f"cached{spec.language_dim}"
    MODEL_CONFIGS["vision"][vkey] = {"name": f"stub/cached-{spec.vision_dim}", "dim": spec.vision_dim}
    MODEL_CONFIGS["language"][lkey] = {"name": f"stub/cached-{spec.language_dim}", "dim": spec.language_dim}
    mm.AutoModelForImageClassification.from_pretrained = staticmethod(lambda *a, **k: _VisionStub())
    mm.AutoModel.from_pretrained = staticmethod(lambda *a, **k: _LanguageStub())
    model = mm.MultimodalRecommender(
        n_users=spec.n_users, n_items=spec.n_items, n_tags=spec.n_tags, num_numerical_features=spec.num_numerical_features,
        embedding_dim=spec.embedding_dim, vision_model_name=vkey if spec.vision_dim else None,
        language_model_name=lkey if spec.language_dim else None, use_contrastive=False,
        num_attention_heads=spec.num_attention_heads, fusion_hidden_dims=list(spec.fusion_hidden_dims),
        fusion_activation=spec.fusion_activation, use_batch_norm=spec.use_batch_norm,
        projection_hidden_dim=spec.projection_hidden_dim, final_activation=spec.final_activation, fusion_type=spec.fusion_type)
    if spec.fusion_type == "attention":      # the documented semantics of the layer (the stock call site is broken)
        model._apply_attention_fusion = types.MethodType(lambda self, feats: self.fusion_layer(feats), model)
    tens = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))) for k, v in sd.items()}
    own = model.state_dict()
    missing, unexpected = model.load_state_dict({k: v for k, v in tens.items() if k in own}, strict=False)
    missing = [k for k in missing if not k.startswith(("vision_model.", "language_model."))]
    if missing:
        raise RuntimeError(f"reference model: missing keys {missing[:4]}")
    return model.eval()


def recommend_block(model, users, feats, k: int, seen_indptr=None, seen_idx=None, chunk_pairs: int = 1 << 18):
    """Every user of `users` against every item through the reference module's forward on flattened (user, item) pairs in
    chunks; seen items dropped, stable descending order, first k (recommender.py:73-106).  -> (idx (n, k), scores (n, k))."""
    import torch
    NI = int(feats["tag_idx"].shape[0])
    items = torch.arange(NI)
    out_i = torch.full((len(users), k), -1, dtype=torch.int64)
    out_s = torch.full((len(users), k), float("-inf"))
    per = max(1, chunk_pairs // NI)
    ones = torch.ones(1, 1, dtype=torch.long)
    with torch.no_grad():
        for u0 in range(0, len(users), per):
            uu = users[u0:u0 + per]
            ui = uu.repeat_interleave(NI)
            ii = items.repeat(len(uu))
            s = model(ui, ii, feats["tag_idx"][ii], image=feats["vis"][ii] if "vis" in feats else None,
                      text_input_ids=feats["txt"][ii] if "txt" in feats else None,
                      text_attention_mask=ones.expand(len(ii), 1) if "txt" in feats else None,
                      numerical_features=feats["num"][ii] if "num" in feats else None)
            s = s.view(len(uu), NI).clone()
            for r, u in enumerate(uu.tolist()):
                if seen_indptr is not None:
                    s[r, seen_idx[int(seen_indptr[u]):int(seen_indptr[u + 1])].long()] = float("-inf")
                order = torch.sort(s[r], descending=True, stable=True).indices[:k]
                keep = order[s[r][order] > float("-inf")]
                out_i[u0 + r, :len(keep)] = keep
                out_s[u0 + r, :len(keep)] = s[r][keep]
    return out_i, out_s


def literal_recommender(model, dataset):
    """The installed `Recommender` on a duck-typed dataset (user_encoder / item_encoder / feature_cache / get_user_history)."""
    import torch
    _, _, Recommender = _import_reference()
    logging.disable(logging.CRITICAL)              # its per-call INFO / ERROR logging is not part of what is timed
    rec = Recommender(model, dataset, torch.device("cpu"))
    rec._debug_has_run_recommender = True          # skip the one-off debug forward (recommender.py:193-219)
    return rec


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))

"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this
container (it cannot travel to the GPU box).  TEST INFRASTRUCTURE ONLY.

Recipe = SURVEY.md §8(c): register cached-feature "backbones" in the mutable
``MODEL_CONFIGS`` registry, patch the two ``from_pretrained`` entry points the
constructor takes for unknown keys to return identity modules whose output has
``pooler_output = input``, and (attention only) route
``_apply_attention_fusion`` to ``self.fusion_layer(list)`` — the layer's
documented semantics (reference ``src/models/layers.py:135-164``).

Run:  python oracle/make_golden.py        (needs /root/reference)
"""
from __future__ import annotations

import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("PXR_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF))

from pixelrec_multimodal_b200 import synthetic as syn  # noqa: E402
from oracle import pxr_oracle as orc  # noqa: E402


class _Out:
    def __init__(self, x):
        self.pooler_output = x
        self.last_hidden_state = None


class _VisionStub(torch.nn.Module):
    def forward(self, pixel_values=None, **kw):
        return _Out(pixel_values)


class _LanguageStub(torch.nn.Module):
    def forward(self, input_ids=None, attention_mask=None, **kw):
        return _Out(input_ids)


def build_reference_model(spec: syn.ModelSpec, sd_np: dict, double: bool):
    import src.models.multimodal as mm
    from src.config import MODEL_CONFIGS

    vkey = f"cached{spec.vision_dim}"
    lkey = f"cached{spec.language_dim}"
    MODEL_CONFIGS["vision"][vkey] = {"name": f"stub/cached-{spec.vision_dim}", "dim": spec.vision_dim}
    MODEL_CONFIGS["language"][lkey] = {"name": f"stub/cached-{spec.language_dim}", "dim": spec.language_dim}
    mm.AutoModelForImageClassification.from_pretrained = staticmethod(lambda *a, **k: _VisionStub())
    mm.AutoModel.from_pretrained = staticmethod(lambda *a, **k: _LanguageStub())
    model = mm.MultimodalRecommender(
        n_users=spec.n_users, n_items=spec.n_items, n_tags=spec.n_tags,
        num_numerical_features=spec.num_numerical_features, embedding_dim=spec.embedding_dim,
        vision_model_name=vkey if spec.vision_dim else None,
        language_model_name=lkey if spec.language_dim else None,
        use_contrastive=False, num_attention_heads=spec.num_attention_heads,
        fusion_hidden_dims=list(spec.fusion_hidden_dims), fusion_activation=spec.fusion_activation,
        use_batch_norm=spec.use_batch_norm, projection_hidden_dim=spec.projection_hidden_dim,
        final_activation=spec.final_activation, fusion_type=spec.fusion_type)
    if spec.fusion_type == "attention":
        model._apply_attention_fusion = types.MethodType(lambda self, feats: self.fusion_layer(feats), model)
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()},
                                                strict=True)
    assert not missing and not unexpected
    model.eval()
    return model.double() if double else model


def spec_cfg(spec: syn.ModelSpec) -> dict:
    return dict(fusion_type=spec.fusion_type, fusion_activation=spec.fusion_activation,
                use_batch_norm=spec.use_batch_norm, final_activation=spec.final_activation,
                num_attention_heads=spec.num_attention_heads)


def run_forward(model, spec, feats, users, items, dtype):
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a)).to(d)
    with torch.no_grad():
        kw = dict(user_idx=t(users, torch.long), item_idx=t(items, torch.long),
                  tag_idx=t(feats["tag_idx"][items], torch.long))
        if spec.vision_dim:
            kw["image"] = t(feats["vis"][items], dtype)
        if spec.language_dim:
            kw["text_input_ids"] = t(feats["txt"][items], dtype)
            kw["text_attention_mask"] = torch.ones(len(items), 1, dtype=torch.long)
        if spec.num_numerical_features:
            kw["numerical_features"] = t(feats["num"][items], dtype)
        return model(**kw)[:, 0].cpu().numpy()


CASES = {
    # name: (ModelSpec kwargs, seed)
    "concat_small": dict(fusion_type="concatenate", embedding_dim=16, vision_dim=32, language_dim=24,
                         fusion_hidden_dims=[64, 32, 16]),
    "gated_small": dict(fusion_type="gated", embedding_dim=16, vision_dim=32, language_dim=24,
                        fusion_hidden_dims=[64, 32, 16]),
    "attention_small": dict(fusion_type="attention", embedding_dim=16, vision_dim=32, language_dim=24,
                            fusion_hidden_dims=[64, 32, 16], num_attention_heads=4),
    "concat_gelu_nobn_tanh": dict(fusion_type="concatenate", embedding_dim=16, vision_dim=32, language_dim=24,
                                  fusion_hidden_dims=[48, 24], fusion_activation="gelu", use_batch_norm=False,
                                  final_activation="tanh"),
    "gated_silu_projhidden": dict(fusion_type="gated", embedding_dim=16, vision_dim=32, language_dim=24,
                                  fusion_hidden_dims=[64, 32, 16], fusion_activation="silu",
                                  projection_hidden_dim=20, final_activation="none"),
    "attention_leaky_2heads": dict(fusion_type="attention", embedding_dim=16, vision_dim=32, language_dim=24,
                                   fusion_hidden_dims=[32], fusion_activation="leaky_relu", num_attention_heads=2),
    "concat_nonum": dict(fusion_type="concatenate", embedding_dim=16, vision_dim=32, language_dim=24,
                         num_numerical_features=0, fusion_hidden_dims=[64, 32, 16]),
    "gated_tanh_act": dict(fusion_type="gated", embedding_dim=16, vision_dim=32, language_dim=24,
                           fusion_hidden_dims=[64, 32, 16], fusion_activation="tanh"),
    # the dims every BASELINE.json config uses (D=64, CLIP-512, SBERT-384, F=7, [512,256,128])
    "concat_full": dict(fusion_type="concatenate"),
    "gated_full": dict(fusion_type="gated"),
    "attention_full": dict(fusion_type="attention"),
}


def main():
    out_dir = REPO / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    summary = {}
    for name, kw in CASES.items():
        full = name.endswith("_full")
        spec = syn.ModelSpec(n_users=24 if full else 12, n_items=64 if full else 40, **kw)
        seed = syn.SEED + (zlib_crc(name) % 1000)
        sd = syn.make_state_dict(spec, seed=seed)
        feats = syn.make_item_features(spec, seed=seed)
        users = np.repeat(np.arange(spec.n_users), spec.n_items).astype(np.int64)
        items = np.tile(np.arange(spec.n_items), spec.n_users).astype(np.int64)
        cal_mean, cal_std = calibrate(sd, spec, feats, users, items)
        m64 = build_reference_model(spec, sd, double=True)
        ref64 = run_forward(m64, spec, feats, users, items, torch.float64)
        m32 = build_reference_model(spec, sd, double=False)
        ref32 = run_forward(m32, spec, feats, users, items, torch.float32)
        mine = orc.forward_pairs(sd, spec_cfg(spec), users, items, feats["tag_idx"][items],
                                 feats["vis"][items] if spec.vision_dim else None,
                                 feats["txt"][items] if spec.language_dim else None,
                                 feats["num"][items] if spec.num_numerical_features else None)
        err = float(np.max(np.abs(mine - ref64)))
        summary[name] = dict(oracle_vs_ref64_maxabs=err, ref32_vs_ref64_maxabs=float(np.max(np.abs(ref32 - ref64))),
                             score_min=float(ref64.min()), score_max=float(ref64.max()))
        assert err < 1e-12, (name, err)
        np.savez_compressed(out_dir / f"forward_{name}.npz",
                            spec=json.dumps(spec.__dict__), seed=seed, cal_mean=cal_mean, cal_std=cal_std,
                            users=users.astype(np.int32), items=items.astype(np.int32),
                            ref64=ref64, ref32=ref32.astype(np.float32))
        print(name, summary[name])

    # ---- Recommender.get_recommendations on a light dataset
    rec_golden = reference_recommender_golden()
    (out_dir / "recommender_lists.json").write_text(json.dumps(rec_golden, indent=1))
    (out_dir / "SUMMARY.json").write_text(json.dumps(summary, indent=1))


def calibrate(sd, spec, feats, users, items):
    """Measure the logit distribution with the oracle, then rescale the output
    layer to (0, 2): returns the (mean, std) that were applied."""
    z = orc.forward_pairs(sd, spec_cfg(spec), users, items, feats["tag_idx"][items],
                          feats["vis"][items] if spec.vision_dim else None,
                          feats["txt"][items] if spec.language_dim else None,
                          feats["num"][items] if spec.num_numerical_features else None, return_logit=True)
    mean, std = float(z.mean()), float(z.std())
    syn.apply_logit_calibration(sd, spec, mean, std)
    return mean, std


def zlib_crc(s):
    import zlib
    return zlib.crc32(s.encode())


class _LightDataset:
    """Just the attributes Recommender touches (src/inference/recommender.py)."""

    def __init__(self, spec, feats, train_indptr, train_idx):
        import pandas as pd
        from sklearn.preprocessing import LabelEncoder
        self.uids, self.iids = syn.user_ids(spec.n_users), syn.item_ids(spec.n_items)
        self.user_encoder = LabelEncoder().fit(self.uids)
        self.item_encoder = LabelEncoder().fit(self.iids)
        self.item_info_df_original = pd.DataFrame({"item_id": self.iids})
        self.feature_cache = {}
        for i, iid in enumerate(self.iids):
            self.feature_cache[iid] = {
                "image": torch.from_numpy(feats["vis"][i]),
                "text_input_ids": torch.from_numpy(feats["txt"][i]),
                "text_attention_mask": torch.ones(1, dtype=torch.long),
                "numerical_features": torch.from_numpy(feats["num"][i]),
                "tag_idx": torch.tensor(int(feats["tag_idx"][i]), dtype=torch.long),
            }
        self._hist = {self.uids[u]: {self.iids[j] for j in train_idx[train_indptr[u]:train_indptr[u + 1]]}
                      for u in range(spec.n_users)}

    def get_user_history(self, user_id):
        return self._hist.get(user_id, set())


def reference_recommender_golden():
    from src.inference.recommender import Recommender
    out = {}
    for ft in ("concatenate", "gated", "attention"):
        spec = syn.ModelSpec(n_users=8, n_items=48, fusion_type=ft, embedding_dim=16, vision_dim=32,
                             language_dim=24, fusion_hidden_dims=[64, 32, 16])
        seed = syn.SEED + 7
        sd = syn.make_state_dict(spec, seed=seed)
        feats = syn.make_item_features(spec, seed=seed)
        uu = np.repeat(np.arange(spec.n_users), spec.n_items).astype(np.int64)
        ii = np.tile(np.arange(spec.n_items), spec.n_users).astype(np.int64)
        cal_mean, cal_std = calibrate(sd, spec, feats, uu, ii)
        indptr, idx, test_item = syn.make_histories(spec.n_users, spec.n_items, seed=seed, lo=3, hi=12)
        model = build_reference_model(spec, sd, double=False)
        ds = _LightDataset(spec, feats, indptr, idx)
        rec = Recommender(model, ds, torch.device("cpu"))
        rec._debug_has_run_recommender = True
        cases = {}
        for u in range(spec.n_users):
            uid = ds.uids[u]
            cases[uid] = {
                "top10_filter": rec.get_recommendations(uid, top_k=10, filter_seen=True),
                "top5_nofilter": rec.get_recommendations(uid, top_k=5, filter_seen=False),
                "cands": rec.get_recommendations(uid, top_k=4, filter_seen=False,
                                                 candidates=[ds.iids[j] for j in (40, 3, 17, 3, 29)] + ["nope"]),
                "score_i5": rec.get_item_score(uid, ds.iids[5]),
            }
        cases["unknown_user"] = rec.get_recommendations("nobody", top_k=5)
        out[ft] = dict(seed=seed, spec=spec.__dict__, cases=cases, cal_mean=cal_mean, cal_std=cal_std,
                       train_indptr=indptr.tolist(), train_idx=idx.tolist(), test_item=test_item.tolist())
    return out


if __name__ == "__main__":
    main()

"""Shared test helpers: rebuild the seeded synthetic weights/features of a golden
case (tests/golden/*.npz store only the reference's OUTPUTS; inputs are
regenerated from (spec, seed) by the version-independent generator)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from pixelrec_multimodal_b200 import synthetic as syn

GOLDEN = Path(__file__).resolve().parent / "golden"

FORWARD_CASES = ["concat_small", "gated_small", "attention_small", "concat_gelu_nobn_tanh",
                 "gated_silu_projhidden", "attention_leaky_2heads", "concat_nonum", "gated_tanh_act",
                 "concat_full", "gated_full", "attention_full"]


def spec_cfg(spec: syn.ModelSpec) -> dict:
    return dict(fusion_type=spec.fusion_type, fusion_activation=spec.fusion_activation,
                use_batch_norm=spec.use_batch_norm, final_activation=spec.final_activation,
                num_attention_heads=spec.num_attention_heads)


def build_case(spec: syn.ModelSpec, seed: int, cal_mean: float, cal_std: float):
    sd = syn.make_state_dict(spec, seed=seed)
    feats = syn.make_item_features(spec, seed=seed)
    syn.apply_logit_calibration(sd, spec, cal_mean, cal_std)
    return sd, feats


def load_forward_case(name: str):
    z = np.load(GOLDEN / f"forward_{name}.npz")
    spec = syn.ModelSpec(**json.loads(str(z["spec"])))
    sd, feats = build_case(spec, int(z["seed"]), float(z["cal_mean"]), float(z["cal_std"]))
    return dict(spec=spec, sd=sd, feats=feats, users=z["users"].astype(np.int64), items=z["items"].astype(np.int64),
                ref64=z["ref64"], ref32=z["ref32"])


def load_recommender_golden():
    return json.loads((GOLDEN / "recommender_lists.json").read_text())


def make_workload(spec: syn.ModelSpec, seed: int, target_std: float = 2.0, sample_users: int = 16,
                  sample_items: int = 256):
    """Seeded weights + features with the output layer calibrated on a small
    oracle sample so scores spread over (0, 1)."""
    from oracle import pxr_oracle as orc
    sd = syn.make_state_dict(spec, seed=seed)
    feats = syn.make_item_features(spec, seed=seed)
    nu, ni = min(sample_users, spec.n_users), min(sample_items, spec.n_items)
    uu = np.repeat(np.arange(nu), ni).astype(np.int64)
    ii = np.tile(np.arange(ni), nu).astype(np.int64)
    z = orc.forward_pairs(sd, spec_cfg(spec), uu, ii, feats["tag_idx"][ii],
                          feats["vis"][ii] if spec.vision_dim else None,
                          feats["txt"][ii] if spec.language_dim else None,
                          feats["num"][ii] if spec.num_numerical_features else None, return_logit=True)
    syn.apply_logit_calibration(sd, spec, float(z.mean()), float(z.std()), target_std)
    return sd, feats


def torch_model_from(spec: syn.ModelSpec, sd, device="cuda", kernel_path="auto", operand_dtype="bf16"):
    import torch
    from pixelrec_multimodal_b200 import FastMultimodalRecommender
    m = FastMultimodalRecommender(
        n_users=spec.n_users, n_items=spec.n_items, n_tags=spec.n_tags,
        num_numerical_features=spec.num_numerical_features, embedding_dim=spec.embedding_dim,
        vision_model_name=f"cached{spec.vision_dim}" if spec.vision_dim else None,
        language_model_name=f"cached{spec.language_dim}" if spec.language_dim else None,
        use_contrastive=False, num_attention_heads=spec.num_attention_heads,
        fusion_hidden_dims=list(spec.fusion_hidden_dims), fusion_activation=spec.fusion_activation,
        use_batch_norm=spec.use_batch_norm, projection_hidden_dim=spec.projection_hidden_dim,
        final_activation=spec.final_activation, fusion_type=spec.fusion_type, kernel_path=kernel_path,
        operand_dtype=operand_dtype)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return m.to(device).eval()


class LightDataset:
    """The attributes the reference Recommender touches on its dataset
    (reference src/inference/recommender.py:58-90, 239-269)."""

    def __init__(self, spec, feats, train_indptr, train_idx):
        import pandas as pd
        import torch
        from sklearn.preprocessing import LabelEncoder
        self.uids, self.iids = syn.user_ids(spec.n_users), syn.item_ids(spec.n_items)
        self.user_encoder = LabelEncoder().fit(self.uids)
        self.item_encoder = LabelEncoder().fit(self.iids)
        self.item_info_df_original = pd.DataFrame({"item_id": self.iids})
        self.feature_cache = {}
        for i, iid in enumerate(self.iids):
            d = {"tag_idx": torch.tensor(int(feats["tag_idx"][i]), dtype=torch.long)}
            if "vis" in feats:
                d["image"] = torch.from_numpy(feats["vis"][i])
            if "txt" in feats:
                d["text_input_ids"] = torch.from_numpy(feats["txt"][i])
                d["text_attention_mask"] = torch.ones(1, dtype=torch.long)
            if "num" in feats:
                d["numerical_features"] = torch.from_numpy(feats["num"][i])
            self.feature_cache[iid] = d
        rows_u, rows_i = [], []
        for u in range(spec.n_users):
            for j in train_idx[train_indptr[u]:train_indptr[u + 1]]:
                rows_u.append(self.uids[u])
                rows_i.append(self.iids[int(j)])
        self.interactions = pd.DataFrame({"user_id": rows_u, "item_id": rows_i})
        self._hist = {self.uids[u]: {self.iids[int(j)] for j in train_idx[train_indptr[u]:train_indptr[u + 1]]}
                      for u in range(spec.n_users)}

    def get_user_history(self, user_id):
        return self._hist.get(user_id, set())

"""Packed feature cache: round trip, per-item dict view, conversion from a
reference-style per-item cache with stub backbones (the same stubs the oracle uses)."""
import numpy as np
import pytest
import torch

from pixelrec_multimodal_b200 import synthetic as syn
from pixelrec_multimodal_b200.packed_cache import PackedFeatureCache, convert_reference_cache, write_packed_cache


def test_round_trip_and_views(tmp_path):
    spec = syn.ModelSpec(n_users=4, n_items=37)
    f = syn.make_item_features(spec, seed=3)
    ids = syn.item_ids(spec.n_items)
    write_packed_cache(tmp_path / "c", ids, f["tag_idx"], f["vis"], f["txt"], f["num"])
    c = PackedFeatureCache(tmp_path / "c")
    assert len(c) == 37 and c.meta["vision_dim"] == spec.vision_dim and ids[5] in c and "nope" not in c
    assert np.array_equal(np.asarray(c.arrays["vis"]), f["vis"]) and np.array_equal(np.asarray(c.tag), f["tag_idx"])
    d = c.get(ids[11])
    assert torch.equal(d["image"], torch.from_numpy(f["vis"][11])) and int(d["tag_idx"]) == int(f["tag_idx"][11])
    assert d["text_input_ids"].dtype == torch.float32 and d["numerical_features"].shape == (spec.num_numerical_features,)
    assert c.get("nope") is None
    order = [ids[i] for i in (30, 2, 2, 17)]
    assert c.rows_for(order).tolist() == [30, 2, 2, 17]
    st = c.to_store("cpu", order=order, chunk_rows=3)
    assert torch.equal(st.vis, torch.from_numpy(f["vis"][[30, 2, 2, 17]])) and st.tag_idx.tolist() == f["tag_idx"][[30, 2, 2, 17]].tolist()
    with pytest.raises(KeyError):
        c.rows_for(["nope"])


def test_missing_modalities_and_empty(tmp_path):
    write_packed_cache(tmp_path / "a", ["x", "y"], [1, 2], vis=np.ones((2, 8)), txt=None, num=None)
    c = PackedFeatureCache(tmp_path / "a")
    assert c.arrays["txt"] is None and "text_input_ids" not in c.get("x") and c.to_store("cpu").txt is None
    write_packed_cache(tmp_path / "e", [], np.zeros(0, np.int64))
    assert len(PackedFeatureCache(tmp_path / "e")) == 0
    with pytest.raises(ValueError):
        write_packed_cache(tmp_path / "bad", ["x"], [1, 2])


def test_convert_reference_cache(tmp_path):
    ref = tmp_path / "vision_stub_lang_stub"
    ref.mkdir()
    ids = [f"i{k}" for k in range(9)]
    g = torch.Generator().manual_seed(0)
    raw = {}
    for it in ids:
        raw[it] = {"image": torch.randn(3, 4, 4, generator=g), "text_input_ids": torch.randint(0, 99, (6,), generator=g),
                   "text_attention_mask": torch.ones(6, dtype=torch.long)}
        torch.save(raw[it], ref / f"{it}.pt")
    enc_img = lambda px: px.flatten(1)[:, :16] * 2.0                 # stub frozen backbones
    enc_txt = lambda tok, mask: (tok.float() * mask.float())[:, :4]
    out = convert_reference_cache(ref, tmp_path / "packed", ids, tag_idx=np.arange(9), num=np.zeros((9, 7)),
                                  encode_image=enc_img, encode_text=enc_txt, batch_size=4)
    c = PackedFeatureCache(out)
    assert c.meta["vision_dim"] == 16 and c.meta["language_dim"] == 4 and c.meta["num_numerical"] == 7
    for r, it in enumerate(ids):
        assert torch.allclose(torch.from_numpy(np.array(c.arrays["vis"][r])), enc_img(raw[it]["image"][None])[0])
        assert torch.allclose(torch.from_numpy(np.array(c.arrays["txt"][r])), enc_txt(raw[it]["text_input_ids"][None], raw[it]["text_attention_mask"][None])[0])
    with pytest.raises(FileNotFoundError):
        convert_reference_cache(ref, tmp_path / "p2", ids + ["ghost"], tag_idx=np.arange(10), encode_image=enc_img)


def test_float16_storage(tmp_path):
    """dtype='float16': vis / txt on disk at half the size, widened to fp32 by get() and to_store(); num stays fp32"""
    spec = syn.ModelSpec(n_users=4, n_items=21)
    f = syn.make_item_features(spec, seed=5)
    ids = syn.item_ids(spec.n_items)
    write_packed_cache(tmp_path / "h", ids, f["tag_idx"], f["vis"], f["txt"], f["num"], dtype="float16")
    write_packed_cache(tmp_path / "s", ids, f["tag_idx"], f["vis"], f["txt"], f["num"])
    assert (tmp_path / "h" / "vis.f16").stat().st_size * 2 == (tmp_path / "s" / "vis.f32").stat().st_size
    assert (tmp_path / "h" / "num.f32").stat().st_size == (tmp_path / "s" / "num.f32").stat().st_size
    c = PackedFeatureCache(tmp_path / "h")
    assert c.dtype == "float16" and PackedFeatureCache(tmp_path / "s").dtype == "float32"
    st = c.to_store("cpu", chunk_rows=8)
    assert st.vis.dtype == torch.float32 and st.txt.dtype == torch.float32
    assert torch.equal(st.vis, torch.from_numpy(f["vis"].astype(np.float16).astype(np.float32)))
    assert torch.equal(st.num, torch.from_numpy(f["num"]))
    assert np.abs(st.vis.numpy() - f["vis"]).max() <= 5e-4 * np.abs(f["vis"]).max() + 1e-6
    d = c.get(ids[3])
    assert d["image"].dtype == torch.float32 and torch.equal(d["image"], st.vis[3])
    with pytest.raises(ValueError):
        write_packed_cache(tmp_path / "x", ids, f["tag_idx"], f["vis"], dtype="int8")

"""Packed, memory-mappable item feature cache.

The reference caches one ``torch.save`` dict per item under
``<cache>/vision_<v>_lang_<l>/<item_id>.pt`` (``src/data/simple_cache.py:51-61,136-175``; writer
``scripts/precompute_cache.py:96-126``) holding backbone *inputs* (pixel tensors, token ids), and re-runs the
frozen backbones for every scored pair.  The full-catalogue path needs the opposite: the *post-backbone*
embedding of every item, once, as dense row-aligned arrays that go to the GPU with a few large copies
(they feed ``pxr_precompute_items`` directly).  Layout of a packed cache directory:

    meta.json      {"version", "n_items", "vision_dim", "language_dim", "num_numerical", "dtype"}
    item_ids.txt   one id per line, row order == item encoder order
    vis.f32 / txt.f32 / num.f32   row-major float32 [n_items, dim]   (absent when dim == 0)
    tag.i64        int64 [n_items]

``dtype="float16"`` stores vis / txt as ``vis.f16`` / ``txt.f16`` (half the bytes on disk and over PCIe; the upload
widens them to the fp32 rows ``pxr_precompute_items`` reads; the numerical features stay fp32).  Opt-in: the cached
embeddings are then rounded to 11 significant bits, which the reference never does.

``convert_reference_cache`` builds it from the reference's per-item files with caller-supplied encoders
(the frozen backbones run once per item); ``PackedFeatureCache`` memory-maps it, serves the reference-style
feature dict per item id (``feature_cache.get(item_id)``, ``recommender.py:239-269``) and uploads
everything as an ``ItemFeatureStore`` through pinned staging buffers.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Callable, Dict, Iterable, Optional, Sequence

import numpy as np
import torch

VERSION = 1
_FILES = {"vis": ("vis.f32", np.float32), "txt": ("txt.f32", np.float32), "num": ("num.f32", np.float32)}


def _file_of(key: str, dtype: str):
    """(file name, numpy dtype) of one feature array under the cache's storage dtype."""
    if dtype == "float16" and key in ("vis", "txt"):
        return _FILES[key][0].replace(".f32", ".f16"), np.float16
    return _FILES[key]


def write_packed_cache(path, item_ids: Sequence[str], tag_idx, vis=None, txt=None, num=None, dtype: str = "float32") -> Path:
    """Write a packed cache; arrays are row-aligned with ``item_ids``.  ``dtype``: storage of vis / txt
    ("float32" | "float16")."""
    if dtype not in ("float32", "float16"):
        raise ValueError("dtype must be 'float32' or 'float16'")
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    n = len(item_ids)
    arrs = {"vis": vis, "txt": txt, "num": num}
    dims = {}
    for k, a in arrs.items():
        if a is None:
            dims[k] = 0
            continue
        fname, ftype = _file_of(k, dtype)
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
        if a.ndim != 2 or a.shape[0] != n:
            raise ValueError(f"{k} must be ({n}, dim), got {a.shape}")
        dims[k] = int(a.shape[1])
        a.astype(ftype, copy=False).tofile(path / fname)
    tag = np.ascontiguousarray(np.asarray(tag_idx, dtype=np.int64))
    if tag.shape != (n,):
        raise ValueError(f"tag_idx must be ({n},), got {tag.shape}")
    tag.tofile(path / "tag.i64")
    (path / "item_ids.txt").write_text("\n".join(str(i) for i in item_ids) + ("\n" if n else ""))
    (path / "meta.json").write_text(json.dumps({"version": VERSION, "n_items": n, "vision_dim": dims["vis"],
                                                "language_dim": dims["txt"], "num_numerical": dims["num"],
                                                "dtype": dtype}, indent=1))
    return path


class PackedFeatureCache:
    """Memory-mapped view of a packed cache directory."""

    def __init__(self, path):
        self.path = Path(path)
        meta = json.loads((self.path / "meta.json").read_text())
        if meta.get("version") != VERSION:
            raise ValueError(f"unsupported packed cache version {meta.get('version')}")
        self.meta = meta
        self.n_items = int(meta["n_items"])
        ids = (self.path / "item_ids.txt").read_text().split("\n")
        self.item_ids = [i for i in ids if i != ""] if self.n_items else []
        if len(self.item_ids) != self.n_items:
            raise ValueError(f"item_ids.txt has {len(self.item_ids)} ids, meta.json says {self.n_items}")
        self._index: Optional[Dict[str, int]] = None
        dims = {"vis": meta["vision_dim"], "txt": meta["language_dim"], "num": meta["num_numerical"]}
        self.arrays = {}
        self.dtype = meta.get("dtype", "float32")
        for k, d in dims.items():
            fname, ftype = _file_of(k, self.dtype)
            self.arrays[k] = (np.memmap(self.path / fname, dtype=ftype, mode="r", shape=(self.n_items, int(d)))
                              if d and self.n_items else None)
        self.tag = np.memmap(self.path / "tag.i64", dtype=np.int64, mode="r", shape=(self.n_items,)) if self.n_items else \
            np.zeros(0, np.int64)

    @property
    def index(self) -> Dict[str, int]:
        if self._index is None:
            self._index = {s: i for i, s in enumerate(self.item_ids)}
        return self._index

    def __len__(self):
        return self.n_items

    def __contains__(self, item_id) -> bool:
        return str(item_id) in self.index

    def get(self, item_id, default=None):
        """Reference-style feature dict of one item (keys of ``src/data/dataset.py:264-303``; ``image`` and
        ``text_input_ids`` carry the cached embeddings)."""
        r = self.index.get(str(item_id))
        if r is None:
            return default
        out = {"tag_idx": torch.tensor(int(self.tag[r]), dtype=torch.long)}
        if self.arrays["vis"] is not None:
            out["image"] = torch.from_numpy(np.array(self.arrays["vis"][r], dtype=np.float32))
        if self.arrays["txt"] is not None:
            out["text_input_ids"] = torch.from_numpy(np.array(self.arrays["txt"][r], dtype=np.float32))
            out["text_attention_mask"] = torch.ones(1, dtype=torch.long)
        if self.arrays["num"] is not None:
            out["numerical_features"] = torch.from_numpy(np.array(self.arrays["num"][r]))
        return out

    def rows_for(self, item_ids: Sequence[str]) -> np.ndarray:
        """Row of every id (KeyError on unknown ids): vectorised replacement of a per-call LabelEncoder.transform."""
        idx = self.index
        return np.fromiter((idx[str(i)] for i in item_ids), dtype=np.int64, count=len(item_ids))

    def to_store(self, device, order: Optional[Sequence[str]] = None, chunk_rows: int = 65536):
        """Upload as an ``ItemFeatureStore`` in ``order`` (default: file order) through pinned staging buffers
        (a few large asynchronous copies instead of one small copy per item)."""
        from .recommender import ItemFeatureStore
        device = torch.device(device)
        rows, absent = None, None
        if order is not None:
            idx = self.index
            rows = np.fromiter((idx.get(str(i), -1) for i in order), dtype=np.int64, count=len(order))
            if (rows < 0).any():
                # ids of the encoder without a cache row: zero features + the per-item 'missing' flag (they score 0.0 like
                # items whose features cannot be fetched in the reference, recommender.py:199-201, 229-230)
                absent = rows < 0
                rows = np.where(absent, 0, rows)
        n = self.n_items if rows is None else len(rows)

        def up(a, dtype):
            if a is None:
                return None
            out = torch.empty((n,) + tuple(a.shape[1:]), dtype=dtype, device=device)
            pin = device.type == "cuda"
            for r0 in range(0, n, chunk_rows):
                r1 = min(n, r0 + chunk_rows)
                blk = np.array(a[r0:r1] if rows is None else a[rows[r0:r1]])      # copy out of the read-only memmap
                if absent is not None:
                    blk[absent[r0:r1]] = 0
                t = torch.from_numpy(blk)
                if pin:
                    t = t.pin_memory()
                out[r0:r1].copy_(t, non_blocking=pin)
            if pin:
                torch.cuda.current_stream(device).synchronize()      # the pinned staging buffers may now be dropped
            return out

        return ItemFeatureStore(up(self.tag, torch.int64), up(self.arrays["vis"], torch.float32),
                                up(self.arrays["txt"], torch.float32), up(self.arrays["num"], torch.float32),
                                None if absent is None else absent.copy())


def convert_reference_cache(ref_cache_dir, out_dir, item_ids: Sequence[str], tag_idx, num=None,
                            encode_image: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                            encode_text: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None,
                            batch_size: int = 64, missing: str = "error") -> Path:
    """One-off conversion of a reference cache directory (``<item_id>.pt`` dicts with ``image``,
    ``text_input_ids``, ``text_attention_mask``; ``simple_cache.py:136-175``) into a packed cache of post-backbone
    embeddings.  ``encode_image(pixel_batch) -> (B, Dv)`` and ``encode_text(ids, mask) -> (B, Dl)`` run the frozen
    backbones ONCE per item (what ``MultimodalRecommender._get_vision_features`` / ``_get_language_features``,
    ``multimodal.py:388-480``, recompute for every scored pair).  ``tag_idx`` / ``num`` come from the processed item
    table (``dataset.py:264-303``), row-aligned with ``item_ids``.  ``missing``: 'error' or 'zeros'."""
    ref = Path(ref_cache_dir)
    vis_rows, txt_rows = [], []
    buf_img, buf_ids, buf_mask = [], [], []

    def flush():
        if buf_img and encode_image is not None:
            with torch.no_grad():
                vis_rows.append(np.asarray(encode_image(torch.stack(buf_img)).float().cpu()))
        if buf_ids and encode_text is not None:
            with torch.no_grad():
                txt_rows.append(np.asarray(encode_text(torch.stack(buf_ids), torch.stack(buf_mask)).float().cpu()))
        buf_img.clear(); buf_ids.clear(); buf_mask.clear()

    template = None
    for it in item_ids:
        f = ref / f"{it}.pt"
        if f.exists():
            d = torch.load(f, map_location="cpu", weights_only=False)
            template = template or d
        elif missing == "zeros" and template is not None:
            d = {k: torch.zeros_like(v) for k, v in template.items() if torch.is_tensor(v)}
        else:
            raise FileNotFoundError(f"no cached features for item {it!r} in {ref}")
        if encode_image is not None:
            buf_img.append(d["image"])
        if encode_text is not None:
            buf_ids.append(d["text_input_ids"]); buf_mask.append(d["text_attention_mask"])
        if max(len(buf_img), len(buf_ids)) >= batch_size:
            flush()
    flush()
    vis = np.concatenate(vis_rows) if vis_rows else None
    txt = np.concatenate(txt_rows) if txt_rows else None
    return write_packed_cache(out_dir, item_ids, tag_idx, vis, txt, num)

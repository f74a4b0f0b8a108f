"""ctypes binding of libpxr.so (include/pxr.h).  There is no CPU fallback: if the
library cannot be loaded, or no CUDA device is present when a handle is created,
the error is raised to the caller."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PXR_MAX_HIDDEN = 8
PXR_MAX_KS = 8
PXR_METRIC_COLS = 9

FUSION = {"concatenate": 0, "gated": 1, "attention": 2}
ACT = {"relu": 0, "gelu": 1, "tanh": 2, "leaky_relu": 3, "silu": 4}
FINAL = {"none": 0, "sigmoid": 1, "tanh": 2}
PATH = {"auto": 0, "simt": 1, "tcgen05": 2}
PRECISION = {"bf16": 0, "fp16": 1}


class PxrError(RuntimeError):
    pass


class PxrConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("fusion", C.c_int32), ("embedding_dim", C.c_int32),
                ("vision_dim", C.c_int32), ("language_dim", C.c_int32), ("num_numerical", C.c_int32),
                ("projection_hidden", C.c_int32), ("n_hidden", C.c_int32),
                ("hidden", C.c_int32 * PXR_MAX_HIDDEN), ("num_heads", C.c_int32), ("activation", C.c_int32),
                ("final_activation", C.c_int32), ("use_batch_norm", C.c_int32), ("n_tags", C.c_int32),
                ("path", C.c_int32), ("precision", C.c_int32)]


_F = C.c_void_p


class PxrWeights(C.Structure):
    _fields_ = ([("struct_size", C.c_int32), ("tag_embedding", _F)] +
                [(n, _F) for n in ("vision_w0", "vision_b0", "vision_w1", "vision_b1",
                                   "language_w0", "language_b0", "language_w1", "language_b1",
                                   "numerical_w0", "numerical_b0", "numerical_w1", "numerical_b1",
                                   "gate_w", "gate_b", "attn_in_w", "attn_in_b", "attn_out_w", "attn_out_b",
                                   "attn_ln_w", "attn_ln_b")] +
                [(n, _F * PXR_MAX_HIDDEN) for n in ("mlp_w", "mlp_b", "bn_w", "bn_b", "bn_mean", "bn_var")] +
                [("out_w", _F), ("out_b", _F), ("bn_eps", C.c_float)])


# every symbol include/pxr.h declares: (restype, argtypes)
_SIGNATURES = {
    "pxr_version": (C.c_int, []),
    "pxr_last_error": (C.c_char_p, [C.c_void_p]),
    "pxr_create": (C.c_int, [C.POINTER(PxrConfig), C.POINTER(C.c_void_p)]),
    "pxr_destroy": (None, [C.c_void_p]),
    "pxr_load_weights": (C.c_int, [C.c_void_p, C.POINTER(PxrWeights), C.c_void_p]),
    "pxr_items_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "pxr_precompute_items": (C.c_int, [C.c_void_p, _F, _F, _F, _F, _F, _F, C.c_int64, C.c_int64, _F, C.c_size_t,
                                       C.c_void_p]),
    "pxr_set_missing_items": (C.c_int, [C.c_void_p, _F, C.c_int64]),
    "pxr_score_topk_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int32]),
    "pxr_score_topk": (C.c_int, [C.c_void_p, _F, _F, C.c_int64, _F, _F, C.c_int32, _F, _F, _F, C.c_size_t,
                                 C.c_void_p]),
    "pxr_set_records_only": (C.c_int, [C.c_void_p, C.c_int]),
    "pxr_rescore_bytes": (C.c_size_t, [C.c_int64]),
    "pxr_rescore_topk": (C.c_int, [C.c_void_p, _F, _F, C.c_int64, _F, C.c_int32, _F, _F, _F, C.c_size_t, C.c_void_p]),
    "pxr_set_small_batch": (C.c_int, [C.c_void_p, C.c_int]),
    "pxr_rescore_lists_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "pxr_rescore_lists": (C.c_int, [C.c_void_p, _F, _F, C.c_int64, _F, C.c_int32, C.c_int32, _F, _F, _F, C.c_size_t, C.c_void_p]),
    "pxr_score_pairs": (C.c_int, [C.c_void_p, _F, _F, _F, C.c_int64, _F, _F, C.c_void_p]),
    "pxr_merge_topk": (C.c_int, [_F, _F, C.c_int32, C.c_int64, C.c_int32, _F, _F, C.c_void_p]),
    "pxr_metrics_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "pxr_metrics": (C.c_int, [_F, C.c_int32, C.c_int64, _F, _F, _F, C.POINTER(C.c_int32), C.c_int32, _F, _F, _F, _F,
                              C.c_size_t, C.c_void_p]),
    "pxr_sample_candidates": (C.c_int, [_F, C.c_int64, _F, _F, C.c_int64, C.c_int32, C.c_uint64, C.c_int32, _F, _F, C.c_void_p]),
    "pxr_weighted_candidates": (C.c_int, [_F, C.c_int64, _F, _F, _F, C.c_int64, C.c_int32, C.c_uint64, C.c_int32, _F, _F, C.c_void_p]),
    "pxr_topk_rows": (C.c_int, [C.c_void_p, _F, C.c_int64, C.c_int64, C.c_int32, _F, _F, C.c_void_p]),
    "pxr_novelty_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "pxr_novelty_metrics": (C.c_int, [_F, C.c_int32, C.c_int64, C.c_int64, _F, _F, _F, _F, _F, _F, C.c_size_t, C.c_void_p]),
    "pxr_gini_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "pxr_gini": (C.c_int, [_F, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _F, _F, C.c_size_t, C.c_void_p]),
    "pxr_ils_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "pxr_intra_list_similarity": (C.c_int, [C.c_void_p, _F, C.c_int32, C.c_int64, _F, C.c_int64, C.c_int32, C.c_int64, _F, _F,
                                            C.c_size_t, C.c_void_p]),
    "pxr_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "pxr_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "pxr_launch_count": (C.c_int64, [C.c_void_p]),
    "pxr_active_path": (C.c_int, [C.c_void_p]),
    "pxr_path_reason": (C.c_char_p, [C.c_void_p]),
    "pxr_set_path": (C.c_int, [C.c_void_p, C.c_int]),
    "pxr_set_rescore": (C.c_int, [C.c_void_p, C.c_int]),
    "pxr_get_rescore": (C.c_int, [C.c_void_p]),
}

_lib = None


def lib_path() -> Path:
    """In-tree libpxr.so; PXR_LIB points at an alternative build (A/B experiments only)."""
    import os
    alt = os.environ.get("PXR_LIB")
    return Path(alt) if alt else Path(__file__).resolve().parent / "libpxr.so"


def load() -> C.CDLL:
    """Load libpxr.so, building it in-tree first when sources are newer (nvcc
    cross-compiles without a GPU).  Raises if neither works."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    path = lib_path()
    try:
        if path == Path(__file__).resolve().parent / "libpxr.so" and _build.needs_build():
            _build.build()
    except Exception as e:  # a prebuilt .so (the GPU box has one) is still usable, but say that it is stale
        if not path.exists():
            raise PxrError(f"libpxr.so is missing and could not be built: {e}") from e
        import warnings
        warnings.warn(f"libpxr.so is older than its sources and the rebuild failed ({e}); loading the existing library",
                      RuntimeWarning, stacklevel=2)
    if not path.exists():
        raise PxrError(f"{path} not found: the CUDA extension is required (no CPU fallback)")
    lib = C.CDLL(str(path))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    return list(_SIGNATURES)

"""CPU checks of the wide gated front end (csrc/score_tc.cu, F_GATEDW).

* The algebra: layer 1 of a gated model is linear in the fused vector and the gate weights sum to 1
  (reference src/models/layers.py:207-223, src/models/multimodal.py:371-386), so
      W1 (g_0 E_u + sum_m g_m f_m) + b1 == g_0 (W1 E_u + b1) + sum_m g_m (W1 f_m + b1)
  -- checked in float64 against the oracle's literal forward, for 4-6 modalities and several embedding dims.
* The chunk-major layout the item GEMM writes and the TMA ring reads (Params::item_q): offsets are unique, inside the
  allocation pxr_tc_item_bytes sizes, 16-byte aligned for the 16-byte shared-memory reads, every (tile, chunk) stage is one
  contiguous block, and the 16 item blocks of a stage start in distinct 16-byte bank groups (conflict-free quarter-warp reads).
* The emulated oracle of the kernel (forward_pairs_lowp) stays within the stated 16-bit band of the exact forward.
"""
import numpy as np
import pytest

from oracle import pxr_oracle as orc
from pixelrec_multimodal_b200 import synthetic as syn
from tests import _cases as cs


def _case(D, **kw):
    spec = syn.ModelSpec(n_users=6, n_items=120, fusion_type="gated", embedding_dim=D, **kw)
    sd = syn.make_state_dict(spec, seed=syn.SEED + 41)
    feats = syn.make_item_features(spec, seed=syn.SEED + 41)
    syn.condition_like_trained(sd, spec, feats)
    return spec, sd, feats


@pytest.mark.parametrize("D,kw", [(16, {}), (128, {}), (192, dict(num_numerical_features=0)), (256, dict(vision_dim=0)),
                                  (128, dict(fusion_hidden_dims=[256, 128, 64]))])
def test_gate_weighted_partials_equal_layer1_of_the_fused_vector(D, kw):
    spec, sd, feats = _case(D, **kw)
    dt = np.float64
    ui = np.repeat(np.arange(spec.n_users), spec.n_items)
    ii = np.tile(np.arange(spec.n_items), spec.n_users)
    g = lambda name: None if feats.get(name) is None else feats[name][ii]
    f = orc.modality_features(sd, "relu", ui, ii, feats["tag_idx"][ii], g("vis"), g("txt"), g("num"), dt)
    ws, bs = orc.fold_batchnorm(sd, True)
    gates = orc.gated_fusion(sd, f, dt, return_gates=True)
    assert np.allclose(gates.sum(axis=1), 1.0, atol=1e-14)
    direct = orc.gated_fusion(sd, f, dt) @ ws[0].T + bs[0]
    split = sum(gates[:, m:m + 1] * (f[m] @ ws[0].T + bs[0]) for m in range(len(f)))
    assert np.max(np.abs(direct - split)) <= 1e-12 * max(1.0, float(np.max(np.abs(direct))))


def q_item_bytes(nm):          # csrc/score_tc.cu: one item's block of a staged chunk, M - 1 rows of 64 columns (16 bit) + padding
    return nm * 128 + 16


def q_stage_bytes(nm):         # one 64-column chunk of one 16-item tile
    return 16 * q_item_bytes(nm)


def q_offset(item, m, col, nm):
    """Byte offset of column `col` of modality m of shard row `item` (the GEMM epilogue of csrc/items_tc.cu, OUT_16 with qt_nm)."""
    return ((item >> 4) * 8 + (col >> 6)) * q_stage_bytes(nm) + (item & 15) * q_item_bytes(nm) + m * 128 + (col & 63) * 2


@pytest.mark.parametrize("nm", [3, 4, 5])
def test_chunk_major_layout_of_the_item_partials(nm):
    n_rows = 77
    rows_padded = (n_rows + 31) // 32 * 32
    total = rows_padded // 16 * 8 * q_stage_bytes(nm)              # pxr_tc_item_bytes (without its 256-byte rounding)
    seen = set()
    for item in range(n_rows):
        for m in range(nm):
            for col in range(0, 512, 8):                           # the epilogue stores 8 columns (16 bytes) at a time
                o = q_offset(item, m, col, nm)
                assert o % 16 == 0 and 0 <= o and o + 16 <= total
                assert o not in seen
                seen.add(o)
    for tile in range(rows_padded // 16):                          # what one TMA bulk copy of the producer fetches
        for c in range(8):
            lo = (tile * 8 + c) * q_stage_bytes(nm)
            offs = [q_offset(tile * 16 + j, m, 64 * c + col, nm) for j in range(16) for m in range(nm) for col in range(0, 64, 8)]
            assert min(offs) == lo and max(offs) + 16 <= lo + q_stage_bytes(nm)
    # shared memory: a quarter warp (8 lanes = 8 items of the tile) reads 16 bytes each at the same (m, column): distinct
    # 16-byte bank groups of the 128-byte bank row
    for j0 in (0, 8):
        groups = {((j0 + l) * q_item_bytes(nm) // 16) % 8 for l in range(8)}
        assert len(groups) == 8
    assert 3 * q_stage_bytes(5) <= 2 * 16 * 1040                    # the three-slot ring fits the concat Pi area (Map::PI_BUF)


@pytest.mark.parametrize("D", [16, 128, 512])
def test_emulated_wide_gated_kernel_is_within_the_16_bit_band(D):
    spec, sd, feats = _case(D)
    NI = spec.n_items
    ii = np.arange(NI)
    for u in (0, 5):
        args = (sd, cs.spec_cfg(spec), np.full(NI, u), ii, feats["tag_idx"], feats["vis"], feats["txt"], feats["num"])
        emu = orc.forward_pairs_lowp(*args)
        ref = orc.forward_pairs(*args)
        assert np.max(np.abs(emu - ref)) <= 3e-2                     # TC_BAND["bf16"] of tests/test_gpu_parity.py

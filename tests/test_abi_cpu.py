"""CPU-only checks of the boundary: the C-ABI library loads without a GPU and exports exactly the symbols
include/dfcsa.h declares; the Python mirror keeps the reference's state_dict layout; the product path refuses CPU
tensors instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dfcsa.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfcsa_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from dfcsa import _lib
    lib = _lib.lib()
    assert lib.dfcsa_version() == 100
    declared = _declared()
    assert declared == sorted(_lib.SYMBOLS), set(declared) ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert isinstance(lib.dfcsa_last_error(), bytes)


def test_bad_arguments_return_error_codes_not_crashes():
    from dfcsa import _lib
    lib = _lib.lib()
    assert lib.dfcsa_conv_gemm(None, 0, None) == 1
    assert b"null params" in lib.dfcsa_last_error()
    assert lib.dfcsa_sgemm(None, None) == 1


def test_struct_layouts_match_header():
    from dfcsa import _lib
    assert ctypes.sizeof(_lib.Seg) == 24
    assert ctypes.sizeof(_lib.ConvParams) == 16 + 3 * 24 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 16 + 16 + 8 + 8
    assert ctypes.sizeof(_lib.ConvEpi) == 8 + 8 + 8 + 8 + 8
    assert ctypes.sizeof(_lib.BnFold) == 5 * 8 + 8 + 8 + 5 * 8 + 8
    assert ctypes.sizeof(_lib.PackJob) == 3 * 8 + 7 * 8 + 4 * 4 + 8
    assert ctypes.sizeof(_lib.WgradParams) == 16 + 32 + 32 + 16 + 8 + 24 + 16 + 8
    assert ctypes.sizeof(_lib.ParamDesc) == 32


def test_module_mirror_has_reference_state_dict_layout():
    from oracle import dfcsa_oracle as O
    from dfcsa.modules import UNetDFCSARes
    m = UNetDFCSARes(3, 1, [64, 128, 256, 512], pool_size=4, ablation_on_qk_channels=8)
    sd, ref = m.state_dict(), O.init_state_dict()
    assert list(sd.keys()) == list(ref.keys()) and len(sd) == 343
    assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in sd)
    assert sum(p.numel() for p in m.parameters()) == 29052083
    assert float(m.down1.res_scale) == pytest.approx(0.1) and float(m.down1.attn_branch[3].gamma) == 0.0


def test_factory_names_and_errors():
    from dfcsa.model_factory import ModelFactory
    cfg = {"model": {"name": "DFC-SA-Res-Block", "features": [8, 16, 32, 64], "pool_size": 4, "ablation_on_qk_channels": 8}}
    m = ModelFactory.get_model(cfg)
    assert type(m).__name__ == "UNetDFCSARes" and m.down1.attn_branch[3].pool_size == 4
    with pytest.raises(ValueError):
        ModelFactory.get_model({"model": {"name": "nope"}})
    with pytest.raises(NotImplementedError):
        ModelFactory.get_model({"model": {"name": "UNet"}})
    # ablations 1(b) / 4: same state-dict keys and shapes as the reference classes (tests/golden/ablations.npz)
    import numpy as np
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ablations.npz"))
    for name in ("UNet_Baseline", "UNet_EncoderOnlyDFC", "UNet_DecoderOnlyDFC", "UNet_BothStandardConv",
             "UNet_AttentionOnly", "UNet_AdditionFusion", "UNet_ConcatFusion"):
        m = ModelFactory.get_model({"model": {"name": name, "features": [8, 8, 16, 16], "pool_size": 4}})
        ref = {k[len(name) + 3:]: z[k] for k in z.files if k.startswith(name + "/w:")}
        sd = m.state_dict()
        assert list(sd.keys()) == list(ref.keys()) and all(tuple(sd[k].shape) == ref[k].shape for k in sd), name


def test_no_cpu_fallback():
    from dfcsa.metrics import calculate_metrics
    from dfcsa.modules import UNetDFCSARes
    m = UNetDFCSARes(3, 1, [8, 16, 32, 64], pool_size=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        calculate_metrics(torch.rand(1, 1, 8, 8), torch.zeros(1, 1, 8, 8), "bce_dice", {})


def test_full_res_attention_model_has_the_reference_state_dict_layout():
    """UNet_FullResAttention shares the key list / shapes of DFC-SA-Res-Block and is reachable
    through the factory name the reference uses (models/model_factory.py:174)."""
    import numpy as np
    from dfcsa.model_factory import ModelFactory
    z = np.load(os.path.join(ROOT, "tests", "golden", "fullres.npz"))
    want = {k[2:]: z[k].shape for k in z.files if k.startswith("w:")}
    m = ModelFactory.get_model({"model": {"name": "UNet_FullResAttention", "features": [8, 8, 16, 16]}})
    sd = m.state_dict()
    assert set(sd.keys()) == set(want.keys())
    assert all(tuple(sd[k].shape) == tuple(want[k]) for k in sd)
    assert m.down1.attn_branch[3].pool_size is None


def test_sliding_window_geometry_matches_the_reference_loop():
    """dfcsa.inference.tile_boxes reproduces the tile coordinates of the reference's predict_large_image
    (inference.py:124-132), including the shifted-back last tiles."""
    from dfcsa.inference import tile_boxes
    for h, w, tile, ov in ((1024, 1024, 224, 50), (500, 700, 224, 50), (224, 224, 224, 50), (300, 230, 224, 0)):
        want = []
        stride = tile - ov
        for y in range(0, h, stride):
            for x in range(0, w, stride):
                y_end, x_end = min(y + tile, h), min(x + tile, w)
                want.append((max(0, y_end - tile), y_end, max(0, x_end - tile), x_end))
        assert tile_boxes(h, w, tile, ov) == want
        cover = torch.zeros(h, w)
        for y0, y1, x0, x1 in want:
            cover[y0:y1, x0:x1] += 1
        assert bool((cover > 0).all())


def test_wgrad_pixel_splits_fill_whole_waves():
    """dfcsa_wgrad_plan (the split choice of the tcgen05 weight-gradient launch, host arithmetic only): the kernel runs one
    CTA per SM, so for every (items, pixel blocks) of the network's shapes the grid must not spill a few CTAs into an
    extra wave (the 2*SMs + 1 = 297-CTA grids of the first build ran three waves), every pixel block must be covered
    exactly once, and no SM may sit idle while the work could be split further."""
    from dfcsa import _lib
    lib = _lib.lib()
    sms = 148
    splits, bps = ctypes.c_int32(), ctypes.c_int64()
    # (items, pixel blocks) of the 224^2 batch-64 training step (profiles/gemm_shapes_r01_p.json) plus edge cases
    cases = [(3, 50176), (24, 3136), (6, 12544), (1, 50176), (96, 784), (2, 12544), (6, 3136), (24, 784), (3, 12544),
             (4, 12544), (96, 196), (16, 784), (32, 196), (8, 784), (1, 16), (4, 16), (16, 16), (1, 1), (400, 7), (149, 1000)]
    for items, pix in cases:
        assert lib.dfcsa_wgrad_plan(ctypes.c_int64(items), ctypes.c_int64(pix), sms, ctypes.byref(splits), ctypes.byref(bps)) == 0
        s, b = splits.value, bps.value
        assert s >= 1 and b >= 1 and (s - 1) * b < pix <= s * b, (items, pix, s, b)       # exact cover, no empty split
        grid = items * s
        waves = -(-grid // sms)
        if items <= sms and pix >= 4 * (sms // items):
            # enough work to fill the machine: the last wave must be (nearly) full - at most 1/8 of the SMs idle in it
            assert grid > (waves - 1) * sms + sms * 7 // 8 or grid == waves * sms or waves * sms - grid < items, (items, pix, s, b, grid)
        # never the "a few CTAs too many" pattern when the item count leaves a choice (items = sms + 1 cannot avoid it)
        if items <= sms // 2:
            assert grid % sms == 0 or grid % sms > sms // 2 or grid < sms, (items, pix, grid)
    assert lib.dfcsa_wgrad_plan(ctypes.c_int64(0), ctypes.c_int64(5), sms, ctypes.byref(splits), ctypes.byref(bps)) == 1

"""GPU parity of the data path (dfcsa_preprocess through dfcsa.data_loader) - bit-exact: against the outputs of the
reference's own transform classes (tests/golden/preprocess_r01.npz), against the oracle on ragged random batches, and
through size-independent properties at dataset-like sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "preprocess_r01.npz")


@pytest.fixture(scope="module")
def D():
    from dfcsa import data_loader
    return data_loader


def test_reference_fixtures_bit_exact(D):
    g = np.load(GOLD)
    for i in range(int(g["n"])):
        oh, ow, seed = (int(v) for v in g[f"meta_{i}"])
        np.random.seed(seed)
        params = D.draw_augmentation(1)
        img, mask = D.preprocess_batch([g[f"img_{i}"]], [g[f"mask_{i}"]], (ow, oh), params)
        assert np.array_equal(img[0].cpu().numpy(), g[f"out_img_{i}_1"]), f"case {i} (train chain)"
        assert np.array_equal(mask[0].cpu().numpy(), g[f"out_mask_{i}_1"])
        img, mask = D.preprocess_batch([g[f"img_{i}"]], [g[f"mask_{i}"]], (ow, oh), None)
        assert np.array_equal(img[0].cpu().numpy(), g[f"out_img_{i}_0"]), f"case {i} (validation chain)"
        assert np.array_equal(mask[0].cpu().numpy(), g[f"out_mask_{i}_0"])


@pytest.mark.parametrize("out_wh", [(32, 32), (48, 40), (224, 224), (99, 17)])
def test_ragged_batch_matches_oracle(D, out_wh):
    """Sources of different sizes (down- and up-scaling, 1-pixel-wide, equal to the output) in ONE call; rotations by
    generic angles, by the exact-transpose angles, with and without the flip."""
    rng = np.random.default_rng(11)
    ow, oh = out_wh
    sizes = [(7, 5), (1, 1), (oh, ow), (300, 217), (64, 500), (33, 2 * ow), (2 * oh + 1, 9), (150, 150), (40, 41)]
    angles = [None, 12.5, -77.25, 90.0, -90.0, 180.0, 0.0, 45.0, 89.999]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    masks = [rng.integers(0, 256, (h, w), dtype=np.uint8) for h, w in sizes]
    params = [(a is not None, a or 0.0, bool(i % 2)) for i, a in enumerate(angles)]
    img, mask = D.preprocess_batch(imgs, masks, (ow, oh), params)
    torch.cuda.synchronize()
    for i in range(len(sizes)):
        ri, rm = P.preprocess_sample(imgs[i], masks[i], ow, oh, *params[i])
        assert np.array_equal(img[i].cpu().numpy(), ri), f"image {i} size {sizes[i]} params {params[i]}"
        assert np.array_equal(mask[i].cpu().numpy(), rm), f"mask {i} size {sizes[i]} params {params[i]}"
    img2, none = D.preprocess_batch(imgs, None, (ow, oh), params)           # images only
    assert none is None and torch.equal(img2, img)


def test_dataset_sized_batch_properties(D):
    """32 photographs' worth of bytes (768x1024 -> 224x224): two samples against the oracle, and for all of them
    flip(preprocess(x)) == preprocess(x, flip) and a quarter turn four times == identity (exact transposes)."""
    rng = np.random.default_rng(3)
    n = 32
    base = rng.integers(0, 256, (n, 768, 1024, 3), dtype=np.uint8)
    imgs = [base[i] for i in range(n)]
    masks = [(rng.random((768, 1024)) > 0.7).astype(np.uint8) * 255 for _ in range(n)]
    plain = [(False, 0.0, False)] * n
    a, am = D.preprocess_batch(imgs, masks, (224, 224), plain)
    b, bm = D.preprocess_batch(imgs, masks, (224, 224), [(False, 0.0, True)] * n)
    assert torch.equal(a.flip(-1), b) and torch.equal(am.flip(-1), bm)
    r90, m90 = D.preprocess_batch(imgs, masks, (224, 224), [(True, 90.0, False)] * n)
    assert torch.equal(torch.rot90(a, 1, (-2, -1)), r90) and torch.equal(torch.rot90(am, 1, (-2, -1)), m90)
    r180, _ = D.preprocess_batch(imgs, masks, (224, 224), [(True, 180.0, False)] * n)
    assert torch.equal(torch.rot90(a, 2, (-2, -1)), r180)
    params = [(True, -33.0, True), (True, 61.5, False)]
    c, cm = D.preprocess_batch(imgs[:2], masks[:2], (224, 224), params)
    for i in range(2):
        ri, rm = P.preprocess_sample(imgs[i], masks[i], 224, 224, *params[i])
        assert np.array_equal(c[i].cpu().numpy(), ri) and np.array_equal(cm[i].cpu().numpy(), rm)
    assert set(torch.unique(am).tolist()) <= {0.0, 1.0}


def test_loader_end_to_end(D, tmp_path):
    Image = pytest.importorskip("PIL.Image")
    (tmp_path / "images").mkdir()
    (tmp_path / "masks").mkdir()
    rng = np.random.default_rng(4)
    src = []
    for i in range(5):
        h, w = int(rng.integers(30, 90)), int(rng.integers(30, 90))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        mask = (rng.random((h, w)) > 0.5).astype(np.uint8) * 255
        Image.fromarray(img).save(tmp_path / "images" / f"s{i}.png")
        Image.fromarray(mask).save(tmp_path / "masks" / f"s{i}.png")
        src.append((img, mask))
    cfg = {"dataset": {"train_dir": str(tmp_path), "val_dir": str(tmp_path), "img_size": [32, 32], "augmentation": True},
           "training": {"batch_size": 2, "num_workers": 2}}
    f = D.DataLoaderFactory(cfg)
    got = list(f.get_val_loader())
    assert [tuple(b["image"].shape) for b in got] == [(2, 3, 32, 32), (2, 3, 32, 32), (1, 3, 32, 32)]
    assert [b["filename"] for b in got] == [["s0.png", "s1.png"], ["s2.png", "s3.png"], ["s4.png"]]
    flat = torch.cat([b["image"] for b in got]).cpu().numpy()
    for i, (img, mask) in enumerate(src):
        assert np.array_equal(flat[i], P.preprocess_sample(img, mask, 32, 32)[0])
    np.random.seed(9)
    tr = list(f.get_train_loader())
    assert sum(b["image"].shape[0] for b in tr) == 5 and all(b["image"].is_cuda and b["mask"].shape[1] == 1 for b in tr)
    # two ranks see disjoint halves of the validation set
    parts = [[n for b in D.DataLoaderFactory(cfg, rank=r, world=2).get_val_loader() for n in b["filename"]] for r in (0, 1)]
    assert sorted(parts[0] + parts[1]) == [f"s{i}.png" for i in range(5)] and not set(parts[0]) & set(parts[1])

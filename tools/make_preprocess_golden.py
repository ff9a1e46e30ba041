#!/usr/bin/env python
"""Generate tests/golden/preprocess_r01.npz by running the REFERENCE's own transform classes
(/root/reference/utils/data_loader.py: ExtCompose[ExtResize, ExtRandomRotation(90), ExtRandomHorizontalFlip, ExtToTensor,
ExtNormalize]) on small synthetic PIL images.  The module is loaded by file path with a stub for its `datasets` import
(that package is not in the reference repository).  Runs only in the build container (needs /root/reference + Pillow +
torchvision); the .npz travels.

Stored per sample: source image / mask (uint8), the numpy seed, the output size, and the reference's outputs (fp32)."""
import importlib.util
import os
import sys
import types

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference():
    stub = types.ModuleType("datasets")
    sub = types.ModuleType("datasets.segmentation_dataset")
    sub.SegmentationDataset = object
    sys.modules["datasets"], sys.modules["datasets.segmentation_dataset"] = stub, sub
    spec = importlib.util.spec_from_file_location("ref_data_loader", "/root/reference/utils/data_loader.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synthetic_pair(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127 + 80 * np.sin(xx / (3.0 + rng.random() * 9)) * np.cos(yy / (2.0 + rng.random() * 7))
    img = np.clip(base[..., None] + rng.normal(0, 30, (h, w, 3)) + rng.integers(-40, 40, (1, 1, 3)), 0, 255).astype(np.uint8)
    blob = ((xx - w * rng.random()) ** 2 + (yy - h * rng.random()) ** 2) < (min(h, w) * (0.2 + 0.3 * rng.random())) ** 2
    mask = np.where(blob, rng.integers(129, 256, (h, w)), rng.integers(0, 127, (h, w))).astype(np.uint8)   # soft values both sides of .5
    return img, mask


def main():
    ref = load_reference()
    rng = np.random.default_rng(7)
    cases = [(41, 57, 32, 32), (90, 64, 32, 32), (33, 33, 48, 40), (120, 75, 32, 32), (32, 32, 32, 32), (25, 70, 40, 48),
             (200, 150, 56, 56), (64, 64, 32, 32)]
    out = {"n": np.int32(len(cases))}
    for i, (h, w, oh, ow) in enumerate(cases):
        img, mask = synthetic_pair(rng, h, w)
        for train in (1, 0):
            tf = ref.ExtCompose([ref.ExtResize((ow, oh)), ref.ExtRandomRotation(degrees=90), ref.ExtRandomHorizontalFlip(),
                                 ref.ExtToTensor(), ref.ExtNormalize()]) if train else \
                ref.ExtCompose([ref.ExtResize((ow, oh)), ref.ExtToTensor(), ref.ExtNormalize()])
            seed = 100 + i
            np.random.seed(seed)
            ti, tm = tf(Image.fromarray(img), Image.fromarray(mask))
            out[f"out_img_{i}_{train}"] = ti.numpy()
            out[f"out_mask_{i}_{train}"] = tm.numpy()
        out[f"img_{i}"], out[f"mask_{i}"] = img, mask
        out[f"meta_{i}"] = np.array([oh, ow, seed], np.int32)
    path = os.path.join(ROOT, "tests", "golden", "preprocess_r01.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

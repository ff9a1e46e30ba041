#!/usr/bin/env python
"""Inference entry point with the reference's command line (inference.py:400-420 of YukiHataRin/DFC-SA-UNet):

  python inference.py --config cfg.yaml --model best_model.pth --input dir [--output results] [--threshold 0.5]
                      [--tile_size 224 --overlap 50] [--resize W H] [--no_slide_window] [--tta]

Loads a reference-format .pth (bare state_dict or the Trainer checkpoint dict), predicts every image under --input
(or --input/original when it exists; --input/mask then gives ground truth for Dice / IoU), writes <stem>_prob.png and
<stem>_pred.png.  Sliding-window tiles and their TTA flips run as batches (dfcsa.inference.predict_large_image).
"""
import argparse
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]

from dfcsa.inference import predict_large_image, predict_single_image, to_normalised_tensor  # noqa: E402
from dfcsa.model_factory import ModelFactory  # noqa: E402


def main(a):
    from PIL import Image
    cfg = yaml.safe_load(open(a.config.replace("\\", "/"), "r", encoding="utf-8"))
    device = torch.device("cuda")
    model = ModelFactory.get_model(cfg)
    sd = torch.load(a.model.replace("\\", "/"), map_location="cpu", weights_only=False)
    model.load_state_dict(sd.get("model_state_dict", sd))            # reference utils/trainer.py:276-296: either form
    model = model.to(device).eval()
    src = os.path.join(a.input, "original") if os.path.isdir(os.path.join(a.input, "original")) else a.input
    gt_dir = os.path.join(a.input, "mask") if os.path.isdir(os.path.join(a.input, "mask")) else None
    os.makedirs(a.output, exist_ok=True)
    rows = []
    for name in sorted(os.listdir(src)):
        stem, ext = os.path.splitext(name)
        if ext.lower() not in (".png", ".jpg", ".jpeg", ".bmp", ".tif", ".tiff"):
            continue
        img = Image.open(os.path.join(src, name)).convert("RGB")
        if a.resize:
            img = img.resize(tuple(a.resize), Image.BILINEAR)
        arr = np.asarray(img)
        if a.no_slide_window:
            h16, w16 = arr.shape[0] // 16 * 16, arr.shape[1] // 16 * 16          # the U-Net needs multiples of 16
            prob = predict_single_image(model, to_normalised_tensor(arr[:h16, :w16], device), device)
        else:
            prob = predict_large_image(model, arr, a.tile_size, a.overlap, device, use_tta=a.tta)
        pred = prob > a.threshold
        Image.fromarray((prob * 255).astype(np.uint8)).save(os.path.join(a.output, f"{stem}_prob.png"))
        Image.fromarray(pred.astype(np.uint8) * 255).save(os.path.join(a.output, f"{stem}_pred.png"))
        if gt_dir:
            cand = [f for f in os.listdir(gt_dir) if os.path.splitext(f)[0] == stem]
            if cand:
                gt = np.asarray(Image.open(os.path.join(gt_dir, cand[0])).convert("L").resize(pred.shape[::-1], Image.NEAREST)) > 127
                gt = gt[:pred.shape[0], :pred.shape[1]]
                tp = float((pred & gt).sum()); fp = float(pred.sum()) - tp; fn = float(gt.sum()) - tp
                rows.append((name, 2 * tp / max(2 * tp + fp + fn, 1e-7), tp / max(tp + fp + fn, 1e-7)))
    if rows:
        csv_dir = a.csv_dir or a.output
        os.makedirs(csv_dir, exist_ok=True)
        with open(os.path.join(csv_dir, "metrics.csv"), "w") as f:
            f.write("file,dice,iou\n")
            for r in rows:
                f.write(f"{r[0]},{r[1]:.6f},{r[2]:.6f}\n")
        print(f"mean dice {np.mean([r[1] for r in rows]):.4f}  mean iou {np.mean([r[2] for r in rows]):.4f}  ({len(rows)} images)")
    print(f"results written to {a.output}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Inference with a trained DFC-SA model (sliding window, TTA, metrics)")
    ap.add_argument("--config", type=str, required=True)
    ap.add_argument("--model", type=str, required=True)
    ap.add_argument("--input", type=str, required=True)
    ap.add_argument("--output", type=str, default="results")
    ap.add_argument("--csv_dir", type=str, default=None)
    ap.add_argument("--threshold", type=float, default=0.5)
    ap.add_argument("--tile_size", type=int, default=224)
    ap.add_argument("--overlap", type=int, default=50)
    ap.add_argument("--resize", nargs=2, type=int, metavar=("WIDTH", "HEIGHT"))
    ap.add_argument("--no_slide_window", action="store_true")
    ap.add_argument("--tta", action="store_true")
    main(ap.parse_args())

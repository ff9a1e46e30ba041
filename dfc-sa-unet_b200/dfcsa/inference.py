"""Inference helpers mirroring the reference's inference.py (predict_single_image :93-102, predict_large_image
:104-153) on the libdfcsa eval path.

The reference predicts a large image tile by tile, batch 1, and with --tta runs three forwards per tile serially.  Here
the tiles (and their TTA flips) are stacked into batches, so a 1024x1024 satellite tile set becomes a handful of
batched eval-mode forwards (BatchNorm folded into per-channel affine, no statistics passes); the sliding-window
geometry, the ImageNet normalisation and the overlap averaging are the reference's.
"""
import numpy as np
import torch

MEAN = (0.485, 0.456, 0.406)      # reference inference.py:64, :121
STD = (0.229, 0.224, 0.225)


def tile_boxes(h, w, tile_size, overlap):
    """(y0, y1, x0, x1) of every sliding-window tile, exactly the reference's loop (inference.py:124-132): stride =
    tile - overlap, the last tile of a row / column is shifted back inside the image."""
    stride = tile_size - overlap
    boxes = []
    for y in range(0, h, stride):
        for x in range(0, w, stride):
            y1, x1 = min(y + tile_size, h), min(x + tile_size, w)
            boxes.append((max(0, y1 - tile_size), y1, max(0, x1 - tile_size), x1))
    return boxes


def to_normalised_tensor(image_u8, device):
    """HxWx3 uint8 RGB -> 1x3xHxW fp32 normalised like transforms.ToTensor() + Normalize(mean, std)."""
    t = torch.as_tensor(np.ascontiguousarray(image_u8), device=device).permute(2, 0, 1).float().div_(255.0)
    mean = torch.tensor(MEAN, device=device).view(3, 1, 1)
    std = torch.tensor(STD, device=device).view(3, 1, 1)
    return ((t - mean) / std).unsqueeze(0)


@torch.no_grad()
def predict_single_image(model, image_tensor, device):
    """reference inference.py:93-102."""
    model.eval()
    out = model(image_tensor.to(device))
    return torch.sigmoid(out).squeeze(0).squeeze(0).cpu().numpy()


@torch.no_grad()
def predict_large_image(model, image, tile_size, overlap, device, use_tta=False, batch_tiles=16):
    """reference inference.py:104-153 with the tiles batched.  image: HxWx3 uint8.  Returns the HxW probability map
    (overlaps averaged).  Tiles smaller than tile_size (image smaller than a tile) go through one by one."""
    model.eval()
    h, w, _ = image.shape
    full = to_normalised_tensor(image, device)[0]                     # 3 x H x W, normalised once
    canvas = torch.zeros(h, w, device=device)
    counts = torch.zeros(h, w, device=device)
    boxes = tile_boxes(h, w, tile_size, overlap)
    groups = {}
    for b in boxes:                                                   # group by tile shape (edge cases of small images)
        groups.setdefault((b[1] - b[0], b[3] - b[2]), []).append(b)
    for (th, tw), bs in groups.items():
        if th < 16 or tw < 16:
            raise ValueError("dfcsa: tile sides must be at least 16 (4 max-pool levels; reference tile_size 224)")
        for i in range(0, len(bs), batch_tiles):
            chunk = bs[i:i + batch_tiles]
            x = torch.stack([full[:, y0:y1, x0:x1] for (y0, y1, x0, x1) in chunk])
            if use_tta:   # original, horizontal flip, vertical flip - averaged after un-flipping (reference :134-141)
                xs = torch.cat([x, torch.flip(x, [3]), torch.flip(x, [2])])
                p = torch.sigmoid(model(xs))
                n = len(chunk)
                p = (p[:n] + torch.flip(p[n:2 * n], [3]) + torch.flip(p[2 * n:], [2])) / 3.0
            else:
                p = torch.sigmoid(model(x))
            for j, (y0, y1, x0, x1) in enumerate(chunk):
                canvas[y0:y1, x0:x1] += p[j, 0]
                counts[y0:y1, x0:x1] += 1
    counts.clamp_(min=1)
    return (canvas / counts).cpu().numpy()

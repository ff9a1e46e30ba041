"""Synthetic segmentation batches for benchmarks and smoke runs (no dataset ships with the reference): a low-frequency
random field gives blob-shaped masks (about a quarter of the pixels positive) and the image is that field plus noise,
so the masks are learnable and the gradients well conditioned.  Host tensors, fp32, NCHW."""
import torch
import torch.nn.functional as F


def synthetic_batch(B, H, W, seed=1, in_channels=3):
    g = torch.Generator().manual_seed(seed)
    low = F.interpolate(torch.randn(B, 1, 7, 7, generator=g), size=(H, W), mode="bicubic", align_corners=False)
    mask = (low > 0.3).float()
    image = 0.5 * torch.randn(B, in_channels, H, W, generator=g) + low
    return image, mask

"""GPU-side data path: drop-in for the reference's utils/data_loader.py (DataLoaderFactory, :76-184) with the
per-sample PIL / torchvision transform chain (:25-74) replaced by one batched CUDA call, `dfcsa_preprocess`
(csrc/preprocess.cu), bit-exact with Pillow's byte arithmetic.

What stays on the host, exactly as in the reference: file decoding (PIL), and the random decisions - drawn per sample
from numpy's global generator in the reference's call order (ExtRandomRotation :40-45: random() < 0.5 then
uniform(-deg, deg); ExtRandomHorizontalFlip :50-53: random() < 0.5), so a run seeded with np.random.seed(k) augments
every sample the way the reference's single-process loader does.  Image.rotate's case analysis (angle -> exact
transpose or inverse affine matrix) is host code in Pillow too and is mirrored by `rotate_plan`.

The reference imports `datasets.segmentation_dataset.SegmentationDataset`, a module its repository does not contain;
`SegmentationDataset` here is the minimal folder dataset that import implies: <root>/images/* with <root>/masks/<same
stem>.* (RGB image, single-channel mask)."""
import ctypes as C
import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib as L

MEAN = (0.485, 0.456, 0.406)      # ExtNormalize defaults, utils/data_loader.py:67
STD = (0.229, 0.224, 0.225)
ROT_NONE, ROT_AFFINE, ROT_90, ROT_180, ROT_270 = 0, 1, 2, 3, 4
_IMG_EXT = (".png", ".jpg", ".jpeg", ".bmp", ".tif", ".tiff")


def rotate_plan(angle, w, h):
    """What PIL.Image.rotate(angle, expand=False) does before it calls into C: multiples of 90 degrees become exact
    transposes (90 / 270 only on square images), anything else an inverse affine matrix about (w/2, h/2) built from
    cos / sin rounded to 15 decimals.  Returns (rot_mode, matrix[6] or None)."""
    angle = angle % 360.0
    if angle == 0:
        return ROT_NONE, None
    if angle == 180:
        return ROT_180, None
    if angle in (90, 270) and w == h:
        return (ROT_90 if angle == 90 else ROT_270), None
    cx, cy = w / 2.0, h / 2.0
    ang = -math.radians(angle)
    m = [round(math.cos(ang), 15), round(math.sin(ang), 15), 0.0, round(-math.sin(ang), 15), round(math.cos(ang), 15), 0.0]
    m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2]
    m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5]
    m[2] += cx
    m[5] += cy
    return ROT_AFFINE, m


def draw_augmentation(n, degrees=90, rng=np.random):
    """[(rotate?, angle, flip?)] for n consecutive samples in the reference's call order on `rng` (numpy's global
    generator by default)."""
    out = []
    for _ in range(n):
        rot, angle = False, 0.0
        if rng.random() < 0.5:
            rot, angle = True, float(rng.uniform(-degrees, degrees))
        out.append((rot, angle, bool(rng.random() < 0.5)))
    return out


def _as_u8_cuda(a, device, ndim):
    t = torch.as_tensor(a)
    if t.dtype != torch.uint8 or t.dim() != ndim:
        raise ValueError(f"dfcsa.data_loader: expected a uint8 array with {ndim} dims, got {t.dtype} {tuple(t.shape)}")
    if not t.is_cuda:
        t = t.contiguous()
        t = (t.pin_memory() if t.numel() >= 1 << 16 else t).to(device, non_blocking=True)
    return t.contiguous()


def preprocess_batch(images, masks, img_size, params=None, mean=MEAN, std=STD, device="cuda"):
    """images: list of uint8 [h, w, 3] arrays / tensors (any sizes); masks: list of uint8 [h, w] or None;
    img_size: (width, height) as in the reference's config (`dataset.img_size`, handed to PIL's resize);
    params: per-sample (rotate?, angle, flip?) or None for the validation chain (resize + tensor + normalise only).
    Returns (image fp32 [n, 3, H, W], mask fp32 [n, 1, H, W] or None) on the device."""
    n = len(images)
    if n == 0:
        raise ValueError("dfcsa.data_loader: empty batch")
    out_w, out_h = int(img_size[0]), int(img_size[1])
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("dfcsa.data_loader: the transform chain runs on the GPU only (no CPU fallback)")
    keep = []
    table = (L.Sample * n)()
    max_h = max_w = 0
    for i in range(n):
        img = _as_u8_cuda(images[i], dev, 3)
        if img.shape[2] != 3:
            raise ValueError("dfcsa.data_loader: images must be RGB ([h, w, 3])")
        h, w = int(img.shape[0]), int(img.shape[1])
        keep.append(img)
        table[i].img, table[i].h, table[i].w = img.data_ptr(), h, w
        if masks is not None:
            m = _as_u8_cuda(masks[i], dev, 2)
            if tuple(m.shape) != (h, w):
                raise ValueError("dfcsa.data_loader: mask and image sizes differ")
            keep.append(m)
            table[i].mask = m.data_ptr()
        rot, angle, flip = params[i] if params is not None else (False, 0.0, False)
        mode, mat = rotate_plan(angle, out_w, out_h) if rot else (ROT_NONE, None)
        table[i].rot_mode, table[i].flip = mode, 1 if flip else 0
        if mat is not None:
            for k in range(6):
                table[i].a[k] = mat[k]
        max_h, max_w = max(max_h, h), max(max_w, w)
    raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
    table_dev = raw.to(dev, non_blocking=False)
    ws_bytes = L.lib().dfcsa_preprocess_workspace_bytes(n, max_h, max_w, out_h, out_w)
    if ws_bytes < 0:
        raise RuntimeError("dfcsa_preprocess_workspace_bytes: bad geometry")
    ws = torch.empty((ws_bytes + 255,), dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    img_out = torch.empty((n, 3, out_h, out_w), dtype=torch.float32, device=dev)
    mask_out = torch.empty((n, 1, out_h, out_w), dtype=torch.float32, device=dev) if masks is not None else None
    mean3, std3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    L.call("dfcsa_preprocess", L.ptr(table_dev), n, max_h, max_w, out_h, out_w, mean3, std3, L.ptr(img_out), L.ptr(mask_out),
           C.c_void_p(ws_ptr), C.c_int64(ws_bytes), L.stream())
    # sources, table and workspace were allocated on the launching stream: the caching allocator keeps them alive (and
    # un-reused) until the four kernels have run
    del keep, table_dev, ws
    return img_out, mask_out


class SegmentationDataset:
    """<root>/images/<name>.<ext> + <root>/masks/<name>.<ext>; items are decoded uint8 arrays (RGB [h, w, 3], L [h, w])."""

    def __init__(self, root, transform=None, img_size=(224, 224)):
        self.root, self.img_size = root, tuple(img_size)
        img_dir, mask_dir = os.path.join(root, "images"), os.path.join(root, "masks")
        if not os.path.isdir(img_dir) or not os.path.isdir(mask_dir):
            raise FileNotFoundError(f"{root}: expected images/ and masks/ sub-directories")
        masks = {os.path.splitext(f)[0]: os.path.join(mask_dir, f) for f in sorted(os.listdir(mask_dir)) if f.lower().endswith(_IMG_EXT)}
        self.items = [(os.path.join(img_dir, f), masks[os.path.splitext(f)[0]]) for f in sorted(os.listdir(img_dir))
                      if f.lower().endswith(_IMG_EXT) and os.path.splitext(f)[0] in masks]
        if not self.items:
            raise FileNotFoundError(f"{root}: no image / mask pairs found")

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        from PIL import Image
        ip, mp = self.items[i]
        with Image.open(ip) as im:
            img = np.array(im.convert("RGB"), dtype=np.uint8)
        with Image.open(mp) as mm:
            mask = np.array(mm.convert("L"), dtype=np.uint8)
        return img, mask

    def filename(self, i):
        return os.path.basename(self.items[i][0])


class GpuLoader:
    """Iterable of {'image': [b, 3, H, W], 'mask': [b, 1, H, W], 'filename': [...]} CUDA batches - the dictionary the
    reference's trainer indexes (utils/trainer.py:116-117, :195-197).  Host threads decode the next batch while the GPU
    transforms the current one.  Under data parallelism every rank takes its own slice of each epoch's order."""

    def __init__(self, dataset, batch_size, img_size, augment, shuffle, num_workers=2, device="cuda", drop_last=False,
                 rank=0, world=1, shard_batches=False):
        self.dataset, self.batch_size, self.img_size = dataset, int(batch_size), tuple(img_size)
        self.augment, self.shuffle, self.device, self.drop_last = augment, shuffle, device, drop_last
        self.num_workers = max(1, int(num_workers))
        self.rank, self.world, self.epoch = int(rank), int(world), 0
        # shard_batches (validation): the batches of the single-process loader are dealt out whole, batch k to rank
        # k % world, so every per-batch metric equals the reference's and their all-reduced mean is the single-process
        # epoch mean; ranks may then see different batch counts, which is fine because the validation loop issues its
        # collectives only after the loop.  Training shards SAMPLES and must give every rank the same number of batches
        # (each train step issues per-bucket all-reduces): see rank_batches().
        self.shard_batches = bool(shard_batches) and self.world > 1

    @staticmethod
    def rank_batches(order, batch_size, rank, world, drop_last, shard_batches=False):
        """The index batches rank `rank` of `world` runs for one epoch's sample `order`.  Sample sharding follows
        torch's DistributedSampler: with drop_last the order is truncated to a multiple of world * batch_size, otherwise
        it is padded (wrapping around) to a multiple of world - either way every rank gets the same number of batches of
        the same sizes, so the per-step collectives can never be left waiting for a rank that ran out of data."""
        bs = int(batch_size)
        if world > 1 and not shard_batches:
            if drop_last:
                order = order[:len(order) // (world * bs) * (world * bs)]
            elif order:
                total = (len(order) + world - 1) // world * world
                order = (order * ((total + len(order) - 1) // len(order)))[:total]
            order = order[rank::world]
        batches = [order[i:i + bs] for i in range(0, len(order), bs)]
        if drop_last and batches and len(batches[-1]) < bs:
            batches.pop()
        if world > 1 and shard_batches:
            batches = batches[rank::world]
        return batches

    def __len__(self):
        return len(self.rank_batches(list(range(len(self.dataset))), self.batch_size, self.rank, self.world, self.drop_last,
                                     self.shard_batches))

    def __iter__(self):
        n = len(self.dataset)
        if self.shuffle and self.world > 1:          # the same permutation on every rank, a different one per epoch
            gen = torch.Generator().manual_seed(1234 + self.epoch)
            order = torch.randperm(n, generator=gen).tolist()
        else:
            order = torch.randperm(n).tolist() if self.shuffle else list(range(n))
        self.epoch += 1
        batches = self.rank_batches(order, self.batch_size, self.rank, self.world, self.drop_last, self.shard_batches)
        with ThreadPoolExecutor(self.num_workers) as pool:
            pending, pending_idx = None, None
            for idx in batches + [None]:
                nxt = [pool.submit(self.dataset.__getitem__, i) for i in idx] if idx is not None else None   # decode ahead
                if pending is not None:
                    pairs = [f.result() for f in pending]
                    params = draw_augmentation(len(pairs)) if self.augment else None
                    img, mask = preprocess_batch([p[0] for p in pairs], [p[1] for p in pairs], self.img_size, params, device=self.device)
                    names = [self.dataset.filename(i) for i in pending_idx] if hasattr(self.dataset, "filename") else [str(i) for i in pending_idx]
                    yield {"image": img, "mask": mask, "filename": names, "index": list(pending_idx)}
                pending, pending_idx = nxt, idx


class DataLoaderFactory:
    """Same constructor and methods as the reference's (utils/data_loader.py:76-184); the loaders yield CUDA batches."""

    def __init__(self, config, device="cuda", rank=0, world=1):
        self.config = config
        self.rank, self.world = rank, world
        self.train_dir = config["dataset"]["train_dir"].replace("\\", "/")
        self.val_dir = config["dataset"]["val_dir"].replace("\\", "/")
        self.batch_size = config["training"]["batch_size"]
        self.num_workers = config["training"]["num_workers"]
        self.img_size = tuple(config["dataset"].get("img_size", [224, 224]))
        self.use_augmentation = config["dataset"]["augmentation"]
        self.device = device

    def get_train_loader(self):
        ds = SegmentationDataset(self.train_dir, img_size=self.img_size)
        return GpuLoader(ds, self.batch_size, self.img_size, augment=bool(self.use_augmentation), shuffle=True,
                         num_workers=self.num_workers, device=self.device, drop_last=self.world > 1, rank=self.rank, world=self.world)

    def get_val_loader(self):
        ds = SegmentationDataset(self.val_dir, img_size=self.img_size)
        # validation: whole batches are dealt out to the ranks and Trainer.validate_epoch all-reduces the sums, so the
        # reported metrics, best_val_loss and the is_best decision are those of the full validation set on every rank
        return GpuLoader(ds, self.batch_size, self.img_size, augment=False, shuffle=False, num_workers=self.num_workers,
                         device=self.device, rank=self.rank, world=self.world, shard_batches=True)

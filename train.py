#!/usr/bin/env python
"""Training entry point with the reference's command line (train.py:98-136 of YukiHataRin/DFC-SA-UNet) on the libdfcsa
kernels:  python train.py --config configs/config_dfc-sa-res-block-p4.yaml [--resume ckpt.pth] [--loss bce_dice] ...

Same YAML schema, same ModelFactory name, same optimizer hyper-parameters and checkpoint dictionary.  Data: the
reference's DataLoaderFactory depends on a `datasets` package that is not part of its repository, so this script takes
image / mask folders in the layout <dir>/images/* and <dir>/masks/* (same file stems; dfcsa.data_loader decodes on host
threads and runs the reference's resize / rotation / flip / normalise chain on the GPU), or --synthetic N to train on N
generated batches (structured blobs, the distribution bench.py uses).  Multi-GPU: launch with torchrun; each rank reads
its share of every batch and gradients are all-reduced per bucket over NCCL (dfcsa.trainer).
"""
import argparse
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]

from dfcsa.model_factory import ModelFactory  # noqa: E402
from dfcsa.optim import FusedSGD  # noqa: E402
from dfcsa.trainer import Trainer  # noqa: E402


def normalize_path(path):
    return path.replace("\\", "/")


class SyntheticBatches:
    """n pre-batched synthetic items (image-correlated blob masks); iterable like a DataLoader."""

    def __init__(self, n, batch, img_size, seed=0):
        from dfcsa.synthetic import synthetic_batch
        H, W = img_size
        self.items = []
        for i in range(n):
            img, msk = synthetic_batch(batch, H, W, seed=seed * 100003 + i)
            self.items.append({"image": img.pin_memory(), "mask": msk.pin_memory()})

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)


def main(config, resume_path=None, synthetic=0):
    if not torch.cuda.is_available():
        raise RuntimeError("dfcsa runs on CUDA (sm_100) only; there is no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    tr_cfg, ds = config["training"], config.get("dataset", {})
    bs, size = int(tr_cfg.get("batch_size", 4)), ds.get("img_size", [224, 224])
    if synthetic:
        train_loader, val_loader = SyntheticBatches(synthetic, bs, size, seed=local), SyntheticBatches(max(1, synthetic // 4), bs, size, seed=1000 + local)
    else:
        # decoded bytes go to the GPU; resize / rotation / flip / normalisation run there, bit-exact with the reference's
        # PIL chain (dfcsa.data_loader, utils/data_loader.py:25-74 of the reference)
        from dfcsa.data_loader import DataLoaderFactory
        config["dataset"].setdefault("augmentation", False)
        config["training"].setdefault("num_workers", 2)
        config["training"].setdefault("batch_size", bs)
        factory = DataLoaderFactory(config, device=device, rank=int(os.environ.get("RANK", "0")), world=world)
        train_loader, val_loader = factory.get_train_loader(), factory.get_val_loader()
    model = ModelFactory.get_model(config).to(device)
    optimizer = FusedSGD(model.parameters(), lr=float(tr_cfg.get("learning_rate", 0.01)), momentum=float(tr_cfg.get("momentum", 0.9)),
                         weight_decay=float(tr_cfg.get("weight_decay", 1e-4)))           # reference train.py:73-78
    trainer = Trainer(model=model, train_loader=train_loader, val_loader=val_loader, optimizer=optimizer, device=device, config=config)
    trainer.train(resume_from=normalize_path(resume_path) if resume_path else None)
    if trainer.rank == 0 and trainer.train_losses:
        print(f"final train loss {trainer.train_losses[-1]:.4f}  dice {trainer.train_dice_scores[-1]:.4f}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Train segmentation model (DFC-SA-Res-Block on libdfcsa)")
    ap.add_argument("--config", type=str, default="configs/config_dfc-sa-res-block-p4.yaml")
    ap.add_argument("--resume", type=str)
    ap.add_argument("--loss", type=str, choices=["dice", "tversky", "bce_dice", "joint"])
    ap.add_argument("--alpha", type=float)
    ap.add_argument("--beta", type=float)
    ap.add_argument("--weight_bce", type=float)
    ap.add_argument("--weight_dice", type=float)
    ap.add_argument("--bce_weight", type=float)
    ap.add_argument("--dice_weight", type=float)
    ap.add_argument("--contour_weight", type=float)
    ap.add_argument("--augmentation", type=lambda x: str(x).lower() == "true")
    ap.add_argument("--synthetic", type=int, default=0, help="train on this many generated batches instead of dataset.train_dir")
    ap.add_argument("--epochs", type=int, help="override training.num_epochs")
    ap.add_argument("--cuda_graph", action="store_true", help="replay each training step from a CUDA graph")
    a = ap.parse_args()
    with open(normalize_path(a.config), "r", encoding="utf-8") as f:
        cfg = yaml.safe_load(f)
    loss = cfg["training"].setdefault("loss", {"type": "dice", "params": {}})
    loss.setdefault("params", {})
    if a.loss is not None:
        loss["type"] = a.loss
    for k in ("alpha", "beta", "weight_bce", "weight_dice", "bce_weight", "dice_weight", "contour_weight"):
        if getattr(a, k) is not None:
            loss["params"][k] = getattr(a, k)
    if a.augmentation is not None:
        cfg.setdefault("dataset", {})["augmentation"] = a.augmentation
    if a.epochs is not None:
        cfg["training"]["num_epochs"] = a.epochs
    if a.cuda_graph:
        cfg["training"]["cuda_graph"] = True
    main(cfg, a.resume, a.synthetic)

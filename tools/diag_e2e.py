"""GPU diagnostic: end-to-end step time with different host->device feeding strategies."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]
import torch
from dfcsa.modules import UNetDFCSARes
from dfcsa.selftest import set_gamma
from dfcsa.trainer import Trainer, device_feeder
from dfcsa import synthetic as O

torch.manual_seed(0)
model = UNetDFCSARes(3, 1, [64, 128, 256, 512], pool_size=4, ablation_on_qk_channels=8)
set_gamma(model, 0.5)
cfg = {"training": {"loss": {"type": "bce_dice", "params": {}}, "num_epochs": 1}, "logging": {"log_dir": "/tmp/x"}}
tr = Trainer(model, None, None, None, "cuda", cfg)
B = 64
host = []
for s in range(2):
    imgs, masks = zip(*[O.synthetic_batch(16, 224, 224, seed=10 * s + j) for j in range(4)])
    host.append((torch.cat(imgs).pin_memory(), torch.cat(masks).pin_memory()))
dev = torch.device("cuda")
devb = [(i.to(dev), m.to(dev)) for i, m in host]
for i in range(4):
    tr.train_step(*devb[i % 2])
torch.cuda.synchronize()
N = 8

def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / N * 1e3:.2f} ms/step", flush=True)

def resident():
    for i in range(N):
        tr.train_step(*devb[i % 2])
def resident_sync():
    for i in range(N):
        r = tr.train_step(*devb[i % 2]); r.stats[:1].cpu().item()
def same_stream():
    for i in range(N):
        img = host[i % 2][0].to(dev, non_blocking=True); msk = host[i % 2][1].to(dev, non_blocking=True)
        r = tr.train_step(img, msk); r.stats[:1].cpu().item()
def feeder():
    for img, msk in device_feeder((host[i % 2] for i in range(N)), dev):
        r = tr.train_step(img, msk); r.stats[:1].cpu().item()
def feeder_nosync():
    for img, msk in device_feeder((host[i % 2] for i in range(N)), dev):
        r = tr.train_step(img, msk)
def cpu_only():
    t0 = time.perf_counter()
    for i in range(N):
        tr.train_step(*devb[i % 2])
    print(f"  host enqueue time: {(time.perf_counter() - t0) / N * 1e3:.2f} ms/step")
for name, fn in (("resident", resident), ("resident+loss readback", resident_sync), ("same-stream H2D", same_stream), ("feeder", feeder),
                 ("feeder, no readback", feeder_nosync), ("resident (host enqueue)", cpu_only), ("feeder again", feeder)):
    timed(name, fn)

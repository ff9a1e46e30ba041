"""GPU diagnostic: how much do two independent training steps overlap when replayed on two streams?  (upper bound on
what running the weight-gradient kernels on a side stream could gain)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]
import torch
from dfcsa.modules import UNetDFCSARes
from dfcsa.selftest import set_gamma
from dfcsa.trainer import Trainer
from dfcsa import synthetic as O

cfg = {"training": {"loss": {"type": "bce_dice", "params": {}}, "num_epochs": 1}, "logging": {"log_dir": "/tmp/x"}}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
trs, data = [], []
for k in range(2):
    torch.manual_seed(k)
    m = UNetDFCSARes(3, 1, [64, 128, 256, 512], pool_size=4, ablation_on_qk_channels=8)
    set_gamma(m, 0.5)
    trs.append(Trainer(m, None, None, None, "cuda", cfg))
    imgs, masks = zip(*[O.synthetic_batch(16, 224, 224, seed=10 * k + j) for j in range(B // 16)])
    data.append((torch.cat(imgs).cuda(), torch.cat(masks).cuda()))
for _ in range(4):
    for k in range(2):
        trs[k].train_step_graphed(*data[k])
torch.cuda.synchronize()
graphs = [list(t._graphs.values())[0][1] for t in trs]
N = 10
def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / N * 1e3:.2f} ms per pair of steps", flush=True)
def serial():
    for _ in range(N):
        graphs[0].replay(); graphs[1].replay()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def parallel():
    for _ in range(N):
        with torch.cuda.stream(s1): graphs[0].replay()
        with torch.cuda.stream(s2): graphs[1].replay()
timed(f"batch {B} serial", serial)
timed(f"batch {B} two streams", parallel)
timed(f"batch {B} serial", serial)
timed(f"batch {B} two streams", parallel)

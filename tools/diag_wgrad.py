"""GPU diagnostic: per-tap / per-channel-half relative error of the tcgen05 weight-gradient kernel."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]
import torch, torch.nn.functional as F
from dfcsa import ops

def run(B, H, W, C, N):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H, W, C, generator=g).cuda().bfloat16()
    dy = torch.randn(B, H, W, N, generator=g).cuda().bfloat16()
    wt = torch.zeros(N, C, 3, 3, device="cuda", requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1)
    (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
    ref = gw.permute(0, 2, 3, 1).reshape(N, 9, C)
    dw = torch.zeros(N, 9 * C, device="cuda")
    ops.conv_wgrad(B, H, W, x.reshape(-1, C), 1, dy.reshape(-1, N), 0, dw, backend=0)
    torch.cuda.synchronize()
    got = dw.view(N, 9, C)
    out = []
    for t in range(9):
        for c0 in range(0, C, 64):
            r, q = ref[:, t, c0:c0 + 64], got[:, t, c0:c0 + 64]
            out.append(f"{float((q - r).norm() / r.norm()):.3f}")
    print(f"B{B} {H}x{W} C{C} N{N}: total {float((got - ref).norm() / ref.norm()):.4f} | per (tap, c-block): {' '.join(out)}")

for cfg in [(1, 32, 32, 64, 64), (1, 32, 32, 128, 64), (1, 28, 28, 128, 64), (1, 16, 4, 128, 64), (1, 32, 32, 128, 128), (1, 32, 32, 256, 64), (2, 14, 14, 64, 128)]:
    try:
        run(*cfg)
    except Exception as e:  # noqa: BLE001
        print(cfg, "FAILED", e)
        break

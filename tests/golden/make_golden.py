"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference (read from /root/reference by
file path; it cannot travel to the GPU box).  Run in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Fixtures (float32 .npz, a few hundred KB in total):
  block_p4.npz   DynamicFusionConvAttnBlock(3->16, pool 4, qk 8) on 2x3x12x12, gamma=0.5: weights, input, output,
                 input-grad-free parameter gradients of sum(out*r)
  net_p4.npz     UNetDFCSARes(features [4,8,16,32], pool 4, qk 4) on 2x3x32x32, gamma=0.5: weights, image, mask,
                 logits, loss / iou / dice from calculate_metrics, all parameter gradients; after one clip+SGD
                 step: BN running stats, the small tensors, and (sum, sum of squares) of every tensor
  net_eval.npz   same weights + the post-step running statistics, eval mode, second input: logits
  fullres.npz    UNet_FullResAttention (ablation 3) with the same weights on 1x3x32x32: logits
  metrics.npz    calculate_metrics('bce_dice') on random probabilities incl. saturated values
  ablations.npz  (python tests/golden/make_golden.py ablations) UNet_Baseline / UNet_EncoderOnlyDFC / UNet_DecoderOnlyDFC
                 / UNet_BothStandardConv / UNet_AttentionOnly / UNet_AdditionFusion / UNet_ConcatFusion (features
                 [8,8,16,16], pool 4) on 2x3x32x32, gamma=0.5: weights, image, mask,
                 logits, bce_dice loss and all parameter gradients, per model
  round2.npz     (python tests/golden/make_golden.py round2) fixtures added in round 2, all from the unmodified reference:
                 lsa16/, lsa64/   LightSelfAttention called on its own (16 ch, pool 4 on 14x14; 64 ch, pool 8 on 20x20), gamma
                                  0.5: x, r, out, d x and parameter gradients of sum(out * r)
                 odd/             UNetDFCSARes([4,8,16,32], pool 4, qk 4) on 2x3x76x92 - NOT multiples of 16, so two of the
                                  four ConvTranspose outputs go through the bilinear re-size of :180-181: logits, loss, grads
                 n64/             the same network on 4x3x64x64 (64 samples per channel in the bottleneck BatchNorms)
                 p16/             pool_size 16 on 2x3x64x64: pooled map larger than the feature map at levels 4-5 (8 -> 16, 4 -> 16)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("DFCSA_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    ref = load("ref_dfc", os.path.join(REF, "models/unet_dfc_sa_res.py"))
    refm = load("ref_metrics", os.path.join(REF, "utils/metrics.py"))
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["refpkg"] = pkg
    load("refpkg.unet_dfc_sa_ablation_branches", os.path.join(REF, "models/unet_dfc_sa_ablation_branches.py"))
    refa = load("refpkg.unet_dfc_sa_ablation_attention", os.path.join(REF, "models/unet_dfc_sa_ablation_attention.py"))
    return ref, refm, refa


def sd_np(sd, prefix="w:"):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def structured(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.nn.functional.interpolate(torch.randn(B, 1, 7, 7, generator=g), size=(H, W), mode="bicubic", align_corners=False)
    return 0.5 * torch.randn(B, 3, H, W, generator=g) + low, (low > 0.3).float()


def ablations():
    ref, refm, refa = load_reference()
    refb = sys.modules["refpkg.unet_dfc_sa_ablation_branches"]
    refp = load("refpkg.unet_dfc_sa_ablation_placement", os.path.join(REF, "models/unet_dfc_sa_ablation_placement.py"))
    reff = load("refpkg.unet_dfc_sa_ablation_fusion", os.path.join(REF, "models/unet_dfc_sa_ablation_fusion.py"))
    img, mask = structured(2, 32, 32, 5)
    d = {"image": img.numpy(), "mask": mask.numpy()}
    for name, ctor in (("UNet_Baseline", lambda: refb.UNet_Baseline(3, 1, [8, 8, 16, 16])),
                       ("UNet_EncoderOnlyDFC", lambda: refp.UNet_EncoderOnlyDFC(3, 1, [8, 8, 16, 16], pool_size=4)),
                       ("UNet_DecoderOnlyDFC", lambda: refp.UNet_DecoderOnlyDFC(3, 1, [8, 8, 16, 16], pool_size=4)),
                       ("UNet_BothStandardConv", lambda: refp.UNet_BothStandardConv(3, 1, [8, 8, 16, 16])),
                       ("UNet_AttentionOnly", lambda: refb.UNet_AttentionOnly(3, 1, [8, 8, 16, 16], pool_size=4)),
                       ("UNet_AdditionFusion", lambda: reff.UNet_AdditionFusion(3, 1, [8, 8, 16, 16], pool_size=4)),
                       ("UNet_ConcatFusion", lambda: reff.UNet_ConcatFusion(3, 1, [8, 8, 16, 16], pool_size=4))):
        torch.manual_seed(0)
        net = ctor()
        with torch.no_grad():
            for n, p in net.named_parameters():
                if n.endswith("gamma"):
                    p.fill_(0.5)
        d.update(sd_np(net.state_dict(), f"{name}/w:"))
        net.train()
        logits = net(img)
        m = refm.calculate_metrics(torch.sigmoid(logits), mask, "bce_dice", {})
        m["loss"].backward()
        d[f"{name}/logits"] = logits.detach().numpy()
        d[f"{name}/loss"] = np.float32(m["loss"].item())
        d.update({f"{name}/g:" + k: p.grad.numpy() for k, p in net.named_parameters()})
    np.savez_compressed(os.path.join(HERE, "ablations.npz"), **d)
    print("ablations.npz written")


def round2():
    ref, refm, refa = load_reference()
    d = {}
    for tag, C, P, hw in (("lsa16", 16, 4, 14), ("lsa64", 64, 8, 20)):
        torch.manual_seed(3)
        att = ref.LightSelfAttention(C, pool_size=P, ablation_on_qk_channels=8)
        with torch.no_grad():
            att.gamma.fill_(0.5)
        x = torch.randn(2, C, hw, hw, requires_grad=True)
        r = torch.randn(2, C, hw, hw)
        out = att(x)
        (out * r).sum().backward()
        d.update(sd_np(att.state_dict(), f"{tag}/w:"))
        d.update({f"{tag}/x": x.detach().numpy(), f"{tag}/r": r.numpy(), f"{tag}/out": out.detach().numpy(), f"{tag}/dx": x.grad.numpy()})
        d.update({f"{tag}/g:" + k: p.grad.numpy() for k, p in att.named_parameters()})
    for tag, P, B, H, W, seed in (("odd", 4, 2, 76, 92, 21), ("n64", 4, 4, 64, 64, 22), ("p16", 16, 2, 64, 64, 23)):
        torch.manual_seed(0)
        net = ref.UNetDFCSARes(3, 1, [4, 8, 16, 32], pool_size=P, ablation_on_qk_channels=4)
        with torch.no_grad():
            for n, p in net.named_parameters():
                if n.endswith("gamma"):
                    p.fill_(0.5)
        d.update(sd_np(net.state_dict(), "net/w:"))        # same seed, and the shapes do not depend on pool_size: stored once
        img, mask = structured(B, H, W, seed)
        net.train()
        logits = net(img)
        m = refm.calculate_metrics(torch.sigmoid(logits), mask, "bce_dice", {})
        m["loss"].backward()
        d.update({f"{tag}/image": img.numpy(), f"{tag}/mask": mask.numpy(), f"{tag}/logits": logits.detach().numpy(),
                  f"{tag}/loss": np.float32(m["loss"].item())})
        d.update({f"{tag}/g:" + k: p.grad.numpy() for k, p in net.named_parameters()})
    np.savez_compressed(os.path.join(HERE, "round2.npz"), **d)
    print("round2.npz written")


def main():
    if sys.argv[1:] == ["ablations"]:
        return ablations()
    if sys.argv[1:] == ["round2"]:
        return round2()
    ref, refm, refa = load_reference()
    torch.manual_seed(0)

    # ---- block ----
    blk = ref.DynamicFusionConvAttnBlock(3, 16, pool_size=4, ablation_on_qk_channels=8)
    with torch.no_grad():
        blk.attn_branch[3].gamma.fill_(0.5)
        for m in blk.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
    w0 = {k: v.clone() for k, v in blk.state_dict().items()}
    x = torch.randn(2, 3, 12, 12)
    r = torch.randn(2, 16, 12, 12)
    blk.train()
    out = blk(x)
    (out * r).sum().backward()
    d = sd_np(w0)
    d.update({"x": x.numpy(), "r": r.numpy(), "out": out.detach().numpy()})
    d.update({"g:" + k: p.grad.numpy() for k, p in blk.named_parameters()})
    d.update(sd_np({k: v for k, v in blk.state_dict().items() if "running" in k or "tracked" in k}, "after:"))
    np.savez_compressed(os.path.join(HERE, "block_p4.npz"), **d)

    # ---- net, train step ----
    net = ref.UNetDFCSARes(3, 1, [4, 8, 16, 32], pool_size=4, ablation_on_qk_channels=4)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.5)
    w0 = {k: v.clone() for k, v in net.state_dict().items()}
    img, mask = structured(2, 32, 32, 1)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    net.train()
    opt.zero_grad()
    logits = net(img)
    m = refm.calculate_metrics(torch.sigmoid(logits), mask, "bce_dice", {"bce_weight": 0.5, "dice_weight": 0.5})
    m["loss"].backward()
    grads = {k: p.grad.clone() for k, p in net.named_parameters()}
    gnorm = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
    opt.step()
    d = sd_np(w0)
    d.update({"image": img.numpy(), "mask": mask.numpy(), "logits": logits.detach().numpy(),
              "loss": np.float32(m["loss"].item()), "iou": np.float64(m["iou"]), "dice": np.float64(m["dice"]),
              "grad_norm": np.float32(gnorm.item())})
    d.update({"g:" + k: v.numpy() for k, v in grads.items()})
    after = net.state_dict()
    d.update(sd_np({k: v for k, v in after.items() if v.numel() < 3000 or "running" in k}, "after:"))
    d.update({"chk:" + k: np.array([v.double().sum().item(), (v.double() ** 2).sum().item()]) for k, v in after.items()})
    np.savez_compressed(os.path.join(HERE, "net_p4.npz"), **d)

    # ---- net, eval mode (uses the running stats left by the step above) ----
    img2, _ = structured(1, 32, 32, 7)
    # eval-mode fixture uses the ORIGINAL weights with the post-step running statistics
    ev = ref.UNetDFCSARes(3, 1, [4, 8, 16, 32], pool_size=4, ablation_on_qk_channels=4)
    esd = {k: (after[k].clone() if ("running" in k or "tracked" in k) else v.clone()) for k, v in w0.items()}
    ev.load_state_dict(esd)
    ev.eval()
    with torch.no_grad():
        lo = ev(img2)
    np.savez_compressed(os.path.join(HERE, "net_eval.npz"), image=img2.numpy(), logits=lo.numpy())

    # ---- ablation 3: full-resolution attention ----
    # NB: the ablation model hard-codes channels // 8 for q/k, so its q/k tensors differ in shape from qk=4 above:
    # store this model's own weights.
    fr = refa.UNet_FullResAttention(3, 1, [8, 8, 16, 16])
    with torch.no_grad():
        for n, p in fr.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.5)
    w0f = {k: v.clone() for k, v in fr.state_dict().items()}
    img3, _ = structured(2, 16, 16, 3)
    fr.train()
    lo = fr(img3)
    d = sd_np(w0f)
    d.update({"image": img3.numpy(), "logits": lo.detach().numpy()})
    np.savez_compressed(os.path.join(HERE, "fullres.npz"), **d)

    # ---- metrics ----
    g = torch.Generator().manual_seed(11)
    p = torch.rand(3, 1, 20, 20, generator=g)
    p[0, 0, 0, :5] = 0.0
    p[0, 0, 1, :5] = 1.0
    t = (torch.rand(3, 1, 20, 20, generator=g) > 0.6).float()
    pr = p.clone().requires_grad_(True)
    m = refm.calculate_metrics(pr, t, "bce_dice", {})
    m["loss"].backward()
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), p=p.numpy(), t=t.numpy(), loss=np.float32(m["loss"].item()),
                        iou=np.float64(m["iou"]), dice=np.float64(m["dice"]), dp=pr.grad.numpy())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Parity report (GPU): logits / loss / gradient errors of the CUDA path vs the oracle for several configurations,
with the per-parameter worst offenders.  Writes gpurun_out/parity_report.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]
from oracle import dfcsa_oracle as O  # noqa: E402  (checker only)
from dfcsa.metrics import calculate_metrics  # noqa: E402
from dfcsa.modules import UNetDFCSARes  # noqa: E402
from dfcsa.selftest import oracle_state, set_gamma  # noqa: E402


def run(features, P, qk, B, H, gamma, seed=0):
    torch.manual_seed(seed)
    model = UNetDFCSARes(3, 1, list(features), pool_size=P, ablation_on_qk_channels=qk)
    set_gamma(model, gamma)
    sd = oracle_state(model)
    img, mask = O.synthetic_batch(B, H, H, seed=1)
    names = O.param_names(sd)
    for k in names:
        sd[k].requires_grad_(True)
    ref_logits = O.unet_forward(img, sd, P, training=True)
    ref_m = O.calculate_metrics(torch.sigmoid(ref_logits), mask, "bce_dice", {})
    ref_g = dict(zip(names, torch.autograd.grad(ref_m["loss"], [sd[k] for k in names])))
    model = model.cuda().train()
    logits = model(img.cuda())
    m = calculate_metrics(torch.sigmoid(logits), mask.cuda(), "bce_dice", {})
    m["loss"].backward()
    torch.cuda.synchronize()
    per = []
    num = den = 0.0
    for n, p in model.named_parameters():
        g, r = p.grad.cpu().double(), ref_g[n].double()
        e2, r2 = float(((g - r) ** 2).sum()), float((r ** 2).sum())
        num += e2; den += r2
        per.append((n, (e2 ** 0.5), (r2 ** 0.5)))
    tot = den ** 0.5
    per.sort(key=lambda t: -t[1])
    res = {"features": list(features), "P": P, "B": B, "H": H, "gamma": gamma,
           "logit_maxabs": float((logits.detach().cpu() - ref_logits.detach()).abs().max()),
           "loss": float(m["loss"].detach()), "loss_ref": float(ref_m["loss"].detach()),
           "grad_rel_l2": (num / den) ** 0.5,
           "worst": [{"param": n, "err_over_total": e / tot, "rel": e / max(r, 1e-30)} for n, e, r in per[:8]]}
    print(json.dumps(res))
    return res


if __name__ == "__main__":
    out = []
    out.append(run((4, 8, 16, 32), 4, 4, 2, 32, 0.5))
    out.append(run((8, 16, 32, 64), 4, 8, 2, 64, 0.5))
    out.append(run((64, 128, 256, 512), 4, 8, 2, 64, 0.5))
    out.append(run((64, 128, 256, 512), 4, 8, 2, 224, 0.5))
    out.append(run((64, 128, 256, 512), 4, 8, 2, 224, 0.0))
    out.append(run((64, 128, 256, 512), 8, 8, 4, 96, 0.5))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)

"""smoke(): one small DFC-SA-Res-Block train step on cuda:0 through the C ABI, checked against the oracle.
The oracle (oracle/dfcsa_oracle.py) is imported here ONLY as the checker."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _oracle():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import dfcsa_oracle
    return dfcsa_oracle


def set_gamma(model, value):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("gamma"):
                p.fill_(value)


def oracle_state(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def grad_rel_l2(model, ref_grads):
    num, den = 0.0, 0.0
    for n, p in model.named_parameters():
        g = p.grad.detach().cpu().double()
        r = ref_grads[n].double()
        num += float(((g - r) ** 2).sum())
        den += float((r ** 2).sum())
    return (num / max(den, 1e-300)) ** 0.5


def forward_backward_parity(features=(64, 128, 256, 512), pool_size=4, qk=8, B=2, H=64, W=64, gamma=0.5, seed=0,
                            full_res_attention=False, model_name=None):
    """Returns dict(logit_maxabs, grad_rel_l2, loss, loss_ref) of dfcsa vs the oracle on identical weights/inputs.
    full_res_attention: the ablation-3 network (UNet_FullResAttention) instead of DFC-SA-Res-Block; model_name: one of
    the other ModelFactory names (ablations 1(b) / 4, which hard-code channels // 8 query/key channels)."""
    O = _oracle()
    from .metrics import calculate_metrics
    from .modules import UNet_FullResAttention, UNetDFCSARes
    torch.manual_seed(seed)
    if model_name is not None:
        from .model_factory import ModelFactory
        model = ModelFactory.get_model({"model": {"name": model_name, "in_channels": 3, "out_channels": 1,
                                                  "features": list(features), "pool_size": pool_size}})
    elif full_res_attention:
        model = UNet_FullResAttention(3, 1, list(features))
    else:
        model = UNetDFCSARes(3, 1, list(features), pool_size=pool_size, ablation_on_qk_channels=qk)
    set_gamma(model, gamma)
    sd = oracle_state(model)
    img, mask = O.synthetic_batch(B, H, W, seed=1)
    # oracle (CPU fp32)
    names = O.param_names(sd)
    for k in names:
        sd[k].requires_grad_(True)
    ref_logits = O.unet_forward(img, sd, pool_size, training=True, full_res_attention=full_res_attention)
    ref_m = O.calculate_metrics(torch.sigmoid(ref_logits), mask, "bce_dice", {})
    ref_grads = dict(zip(names, torch.autograd.grad(ref_m["loss"], [sd[k] for k in names])))
    # dfcsa (CUDA)
    model = model.cuda().train()
    logits = model(img.cuda())
    m = calculate_metrics(torch.sigmoid(logits), mask.cuda(), "bce_dice", {})
    m["loss"].backward()
    torch.cuda.synchronize()
    return {
        "logit_maxabs": float((logits.detach().cpu() - ref_logits.detach()).abs().max()),
        "grad_rel_l2": grad_rel_l2(model, ref_grads),
        "loss": float(m["loss"]), "loss_ref": float(ref_m["loss"]),
        "dice": m["dice"], "dice_ref": ref_m["dice"],
    }


def eval_parity(features=(64, 128, 256, 512), pool_size=4, qk=8, B=2, H=224, W=224, gamma=0.5, seed=0):
    """Eval-mode (running-statistics BatchNorm, reference inference.py:100) logits of dfcsa vs the oracle.  The running
    statistics are a realistic trained state: exactly the batch statistics of one synthetic batch (one oracle training
    forward from the (0, 1) initial buffers, momentum step undone); the compared forward runs on a DIFFERENT batch."""
    O = _oracle()
    from .modules import UNetDFCSARes
    torch.manual_seed(seed)
    model = UNetDFCSARes(3, 1, list(features), pool_size=pool_size, ablation_on_qk_channels=qk)
    set_gamma(model, gamma)
    sd = oracle_state(model)
    img, _ = O.synthetic_batch(B, H, W, seed=1)
    with torch.no_grad():
        O.unet_forward(img, sd, pool_size, training=True)
        for k in sd:
            if k.endswith("running_mean"):
                sd[k] = sd[k] / 0.1
            elif k.endswith("running_var"):
                sd[k] = (sd[k] - 0.9) / 0.1
        img2, _ = O.synthetic_batch(B, H, W, seed=2)
        ref = O.unet_forward(img2, sd, pool_size, training=False)
        model.load_state_dict(sd)
        model = model.cuda().eval()
        logits = model(img2.cuda())
        torch.cuda.synchronize()
    return {"logit_maxabs": float((logits.cpu() - ref).abs().max()), "logit_ref_absmax": float(ref.abs().max())}


def smoke():
    if not torch.cuda.is_available():
        raise RuntimeError("dfcsa smoke() needs a CUDA device")
    from . import _lib
    if not _lib.lib().dfcsa_device_ok():
        raise RuntimeError("dfcsa smoke(): cuda:0 is not an sm_100 device")
    torch.cuda.set_device(0)
    r = forward_backward_parity(B=2, H=64, W=64, gamma=0.5)
    print("dfcsa smoke:", r)
    assert r["logit_maxabs"] <= 2e-2, r
    assert r["grad_rel_l2"] <= 3e-2, r
    assert abs(r["loss"] - r["loss_ref"]) <= 2e-3, r

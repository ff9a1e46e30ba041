#!/usr/bin/env python
"""Timings of the other BASELINE.json configurations on one B200 (bench.py is the headline metric only):

  c1  DFC-SA-Res-Block P4, 224^2, batch 4, full train step (the reference's own CPU-runnable case) - latency bound
  c2  pool-size sweep P4 / P16 / P32, 224^2, batch 64, train
  c3  ablation 3 (UNet_FullResAttention), 224^2, batch 1, train
  c4  512^2, 32 images per GPU (the per-GPU share of global batch 256 on 8 GPUs), train, + peak memory
  c5  eval-mode batched inference at 1024^2, batch 1/2/4/8: images/s and p50 latency
  c6  data path: 64 decoded 768x1024 photographs -> augmented, normalised 224^2 tensors on the GPU
      (dfcsa_preprocess; sources resident in HBM, and end to end from pinned host memory) beside the same chain run by
      Pillow + numpy on one host core, which is what each of the reference's DataLoader workers does

  python tools/bench_configs.py [c1 c2 ...] --out gpurun_out/configs.json
All numbers are CUDA-event timings after warm-up, inputs resident on the device, synthetic structured images.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]

import torch  # noqa: E402

from dfcsa.modules import UNet_FullResAttention, UNetDFCSARes  # noqa: E402
from dfcsa.selftest import set_gamma  # noqa: E402
from dfcsa.trainer import Trainer  # noqa: E402
from dfcsa import synthetic as O  # noqa: E402

CFG = {"training": {"loss": {"type": "bce_dice", "params": {}}, "num_epochs": 1}, "logging": {"log_dir": "/tmp/dfcsa_cfg"}}
FEATURES = [64, 128, 256, 512]
PROFILE = False


def batch(B, hw, seed=1):
    parts = [O.synthetic_batch(min(8, B - i), hw, hw, seed=seed + i) for i in range(0, B, 8)]
    return torch.cat([p[0] for p in parts]).cuda(), torch.cat([p[1] for p in parts]).cuda()


def time_steps(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    return {"ms_mean": sum(ms) / len(ms), "ms_p50": ms[len(ms) // 2], "ms_min": ms[0]}


def train_case(model, B, hw, warmup=3, steps=8, graph=False):
    torch.cuda.reset_peak_memory_stats()
    tr = Trainer(model, None, None, None, "cuda", CFG)
    img, mask = batch(B, hw)
    step = (lambda: tr.train_step_graphed(img, mask)) if graph else (lambda: tr.train_step(img, mask))
    t0 = time.time()
    r = time_steps(step, warmup, steps)
    if PROFILE:
        from dfcsa import _lib
        _lib.PROF = prof = _lib.Profiler()
        tr.train_step(img, mask)
        _lib.PROF = None
        summ = sorted(prof.summary().items(), key=lambda kv: -kv[1]["ms"])
        r["profile_ms"] = {k: round(v["ms"], 3) for k, v in summ[:14]}
        det = sorted(prof.detail(tags=("bgemm_tc",)).items(), key=lambda kv: -kv[1]["ms"])
        r["bgemm_ms"] = {k: [round(v["ms"], 3), v["launches"]] for k, v in det[:12]}
    r.update({"batch": B, "hw": hw, "img_per_s": B / (r["ms_mean"] * 1e-3), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
              "wall_s": time.time() - t0, "cuda_graph": graph})
    del tr
    torch.cuda.empty_cache()
    return r


def new_model(P=4, full_res=False):
    torch.manual_seed(0)
    m = UNet_FullResAttention(3, 1, FEATURES) if full_res else UNetDFCSARes(3, 1, FEATURES, pool_size=P, ablation_on_qk_channels=8)
    set_gamma(m, 0.5)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--profile", action="store_true", help="add per-entry-point CUDA-event times of one extra step")
    args = ap.parse_args()
    global PROFILE
    PROFILE = args.profile
    res = {}

    def run(name, fn):
        try:
            res[name] = fn()
        except Exception as e:  # noqa: BLE001
            res[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            torch.cuda.empty_cache()
        print(name, json.dumps(res[name]), flush=True)
        json.dump(res, open(args.out, "w"), indent=1)

    if "c1" in args.cases:
        run("c1_p4_224_b4_train", lambda: train_case(new_model(4), 4, 224, steps=20))
        if hasattr(Trainer, "train_step_graphed"):
            run("c1_p4_224_b4_train_cudagraph", lambda: train_case(new_model(4), 4, 224, steps=20, graph=True))
    if "c2" in args.cases:
        for P in (4, 16, 32) if not os.environ.get("C2_P") else [int(os.environ["C2_P"])]:
            run(f"c2_p{P}_224_b64_train", lambda P=P: train_case(new_model(P), 64, 224, steps=5))
    if "c3" in args.cases:
        run("c3_fullres_224_b1_train", lambda: train_case(new_model(full_res=True), 1, 224, warmup=1, steps=2))
    if "c4" in args.cases:
        run("c4_p4_512_b32_train", lambda: train_case(new_model(4), 32, 512, warmup=2, steps=4))
    if "c5" in args.cases:
        model = new_model(4).cuda().eval()
        for graph in (False, True):
            model.eval_cuda_graph = graph
            for B in (1, 2, 4, 8):
                def infer(B=B):
                    img, _ = batch(B, 1024)
                    with torch.no_grad():
                        r = time_steps(lambda: model(img), 10, 50)
                    r.update({"batch": B, "hw": 1024, "img_per_s": B / (r["ms_mean"] * 1e-3), "p50_latency_ms": r["ms_p50"], "cuda_graph": graph})
                    if PROFILE and not graph:
                        from dfcsa import _lib
                        _lib.PROF = prof = _lib.Profiler()
                        with torch.no_grad():
                            model(img)
                        _lib.PROF = None
                        summ = sorted(prof.summary().items(), key=lambda kv: -kv[1]["ms"])
                        r["profile_ms"] = {k: [round(v["ms"], 3), v["launches"]] for k, v in summ[:16]}
                        det = sorted(prof.detail(tags=("conv_tc", "conv_simt")).items(), key=lambda kv: -kv[1]["ms"])
                        r["gemm_ms"] = {k: [round(v["ms"], 3), round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)] for k, v in det[:24]}
                    return r
                run(f"c5_p4_1024_b{B}_eval" + ("_cudagraph" if graph else ""), infer)
    if "c6" in args.cases:
        run("c6_datapath_768x1024_to_224_b64", datapath_case)
    print(json.dumps(res))


def datapath_case(n=64, src_hw=(768, 1024), out=224):
    import numpy as np
    from dfcsa import data_loader as D
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (*src_hw, 3), dtype=np.uint8) for _ in range(n)]
    masks = [(rng.random(src_hw) > 0.7).astype(np.uint8) * 255 for _ in range(n)]
    np.random.seed(0)
    params = D.draw_augmentation(n)
    dimgs = [torch.from_numpy(a).cuda() for a in imgs]
    dmasks = [torch.from_numpy(a).cuda() for a in masks]
    r = time_steps(lambda: D.preprocess_batch(dimgs, dmasks, (out, out), params), 3, 10)
    res = {"n": n, "src_hw": list(src_hw), "out": out, "gpu_ms_resident": r["ms_mean"], "gpu_img_per_s_resident": n / (r["ms_mean"] * 1e-3),
           "src_bytes_per_batch": int(sum(a.nbytes for a in imgs) + sum(a.nbytes for a in masks))}
    res["gpu_gb_per_s_resident"] = (res["src_bytes_per_batch"] + n * 4 * out * out * 4) / (r["ms_mean"] * 1e-3) / 1e9
    pimgs = [torch.from_numpy(a).pin_memory() for a in imgs]
    pmasks = [torch.from_numpy(a).pin_memory() for a in masks]
    r = time_steps(lambda: D.preprocess_batch(pimgs, pmasks, (out, out), params), 2, 5)
    res.update({"gpu_ms_from_host": r["ms_mean"], "gpu_img_per_s_from_host": n / (r["ms_mean"] * 1e-3)})
    try:
        from PIL import Image
        mean, std = np.asarray(D.MEAN, np.float32)[:, None, None], np.asarray(D.STD, np.float32)[:, None, None]
        k = 16
        t0 = time.time()
        for i in range(k):
            im, mk = Image.fromarray(imgs[i]), Image.fromarray(masks[i])
            im, mk = im.resize((out, out), Image.BILINEAR), mk.resize((out, out), Image.NEAREST)
            rot, ang, flip = params[i]
            if rot:
                im, mk = im.rotate(ang, Image.BILINEAR), mk.rotate(ang, Image.NEAREST)
            if flip:
                im, mk = im.transpose(Image.FLIP_LEFT_RIGHT), mk.transpose(Image.FLIP_LEFT_RIGHT)
            x = (np.asarray(im, np.float32).transpose(2, 0, 1) / np.float32(255) - mean) / std
            m = ((np.asarray(mk, np.float32) / 255.0) > 0.5).astype(np.float32)
        dt = time.time() - t0
        res.update({"pillow_one_core_img_per_s": k / dt, "pillow_sample": f"{k} images, 1 thread"})
    except ImportError:
        res["pillow_one_core_img_per_s"] = None
    return res


if __name__ == "__main__":
    main()

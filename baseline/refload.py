"""Loader for the UNMODIFIED reference classes (bench.py's `--impl reference` / `--impl reference-gpu` arms only).

The reference (YukiHataRin/DFC-SA-UNet) has no setup.py / pyproject and its package __init__ files import modules that
are not in the repository, so it cannot be pip-installed; `__graft_entry__.build()` instead copies the
four files of the hot path byte for byte from /root/reference into the git-ignored baseline/_ref/ (which travels to the
GPU box with the repo snapshot), and this module loads them by file path exactly as SURVEY.md App. E describes.  Nothing
under dfc-sa-unet_b200/ imports this file.
"""
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("models/unet_dfc_sa_res.py", "models/unet_dfc_sa_ablation_attention.py", "models/unet_dfc_sa_ablation_branches.py",
         "utils/metrics.py")


def install(src_root="/root/reference"):
    """Copy the reference's hot-path files into baseline/_ref/ (no-op when the reference tree is absent)."""
    if not os.path.isdir(src_root):
        return False
    for rel in FILES:
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, rel), dst)
    return True


def available():
    return all(os.path.exists(os.path.join(REF_DIR, rel)) for rel in FILES)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """-> (models module with UNetDFCSARes, metrics module with calculate_metrics, ablation-attention module)."""
    if not available():
        raise FileNotFoundError(f"{REF_DIR} does not hold the reference files; run __graft_entry__.build() where /root/reference exists")
    sys.dont_write_bytecode = True
    ref = _load("dfcsa_ref_models", os.path.join(REF_DIR, "models/unet_dfc_sa_res.py"))
    refm = _load("dfcsa_ref_metrics", os.path.join(REF_DIR, "utils/metrics.py"))
    pkg = types.ModuleType("dfcsa_refpkg")
    pkg.__path__ = [os.path.join(REF_DIR, "models")]
    sys.modules["dfcsa_refpkg"] = pkg
    _load("dfcsa_refpkg.unet_dfc_sa_ablation_branches", os.path.join(REF_DIR, "models/unet_dfc_sa_ablation_branches.py"))
    refa = _load("dfcsa_refpkg.unet_dfc_sa_ablation_attention", os.path.join(REF_DIR, "models/unet_dfc_sa_ablation_attention.py"))
    return ref, refm, refa

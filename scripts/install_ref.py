#!/usr/bin/env python
"""Install the UNMODIFIED reference (gaspardol/MonteCarloPredictiveCoding) into git-ignored ``baseline/_ref/`` so that
``bench.py --impl reference`` and the cpu_baseline leg can time the reference's own PyTorch path on the GPU box's host
cores (``/root/reference`` does not exist there; ``baseline/_ref`` travels with the gpurun snapshot).

The reference is a flat directory of scripts (no setup.py / pyproject.toml), so ``pip install --target`` has nothing to
build; this script copies the files the hot path needs, byte for byte:
    predictive_coding/{__init__,pc_trainer,pc_layer,utils}.py      the library (SURVEY §2, rows marked with a star)
    utils/{__init__,model,training_evaluation,data}.py             get_model / random_step / trainer factories
Nothing under baseline/_ref is ever imported by the product package or committed to git.

    python scripts/install_ref.py [--ref /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = [
    "predictive_coding/__init__.py", "predictive_coding/pc_trainer.py", "predictive_coding/pc_layer.py",
    "predictive_coding/utils.py",
    "utils/__init__.py", "utils/model.py", "utils/training_evaluation.py", "utils/data.py",
    "requirements.txt", "README.md",
]


def install(ref="/root/reference", dest=None, quiet=False):
    dest = dest or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(ref):
        raise FileNotFoundError(f"reference checkout not found at {ref}")
    manifest = {}
    for rel in FILES:
        src = os.path.join(ref, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref, "files": manifest}, fh, indent=1)
    if not quiet:
        print(f"installed {len(manifest)} reference files into {dest}")
    return dest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("MCPC_REFERENCE", "/root/reference"))
    ap.add_argument("--dest", default=None)
    a = ap.parse_args()
    install(a.ref, a.dest)
    sys.exit(0)

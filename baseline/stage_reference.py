"""Stages the UNMODIFIED reference hot-path modules for the benchmark's reference arms.

    python baseline/stage_reference.py [--src /root/reference]

The reference (DveloperY0115/torch-NeRF) is pure Python with no setup.py / pyproject.toml, so there is nothing to
`pip install`: this script copies the package directories the hot path imports --
`torch_nerf/src/{renderer,scene,network,signal_encoder}` plus the two `__init__.py` above them -- byte for byte into
`baseline/_ref/` (git-ignored, NOT gpurun-ignored, so it travels to the GPU box; no reference source enters the
repository's history) and writes `baseline/_ref/MANIFEST.json` with the sha256 of every file, which
`baseline/ref_harness.py` re-checks before it imports anything.  Runs in the dev container only (`/root/reference`
does not exist on the GPU box); `__graft_entry__.build()` calls it when the reference tree is present.
"""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PACKAGES = ["renderer", "scene", "network", "signal_encoder"]


def sha256(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def stage(src_root="/root/reference", verbose=True):
    src_pkg = os.path.join(src_root, "torch_nerf")
    if not os.path.isdir(os.path.join(src_pkg, "src")):
        raise FileNotFoundError(f"{src_pkg}/src not found: the reference tree is only available in the dev container")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    files = [("torch_nerf/__init__.py", os.path.join(src_pkg, "__init__.py")),
             ("torch_nerf/src/__init__.py", os.path.join(src_pkg, "src", "__init__.py"))]
    for pkg in PACKAGES:
        for dirpath, _, names in os.walk(os.path.join(src_pkg, "src", pkg)):
            for name in sorted(names):
                if name.endswith(".py"):
                    full = os.path.join(dirpath, name)
                    files.append((os.path.relpath(full, src_root), full))
    manifest = {}
    for rel, full in files:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(full, dst)
        manifest[rel] = sha256(dst)
        assert manifest[rel] == sha256(full)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "DveloperY0115/torch-NeRF (unmodified copies)", "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest)} reference files into {DEST}")
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    stage(ap.parse_args().src)

"""Builds libnerf_b200.so (the C-ABI library) and libnerf_b200_selftest.so (tcgen05 building-block checks and
micro-benchmarks for tests/ and tools/, see include/nerf_b200_debug.h) in-tree with nvcc for sm_100a.

    python torch-nerf_b200/build.py            # or  __graft_entry__.build()

No torch involvement: plain `nvcc -shared`.  The built .so is git-ignored but travels to the GPU box.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
# experiments: NERF_B200_LIB_SUFFIX=_x NERF_B200_NVCC_EXTRA="-DFOO=1" builds lib/libnerf_b200_x.so next to the product library
_SUFFIX = os.environ.get("NERF_B200_LIB_SUFFIX", "")
LIB = os.path.join(LIB_DIR, f"libnerf_b200{_SUFFIX}.so")
LIB_SELFTEST = os.path.join(LIB_DIR, f"libnerf_b200{_SUFFIX}_selftest.so")
SOURCES = ["api.cu", "rays.cu", "encode_composite.cu", "mlp_f32.cu", "mlp_tc_pack.cu", "mlp_tc_fwd.cu", "mlp_tc_bwd.cu",
           "optim.cu", "dp_exchange.cu"]
SELFTEST_SOURCES = ["mlp_tc_selftest.cu"]  # links against libnerf_b200.so (error string, SM count)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
] + os.environ.get("NERF_B200_NVCC_EXTRA", "").split()


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libnerf_b200.so cannot be built")


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files += [os.path.join(ROOT, "include", h) for h in ("nerf_b200.h", "nerf_b200_debug.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())
    h.update(" ".join(a for a in NVCC_FLAGS if not a.startswith("-I")).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, f"build{_SUFFIX}.sha256")
    digest = _digest()
    if (not force and os.path.exists(LIB) and os.path.exists(LIB_SELFTEST) and os.path.exists(stamp)
            and open(stamp).read().strip() == digest):
        return LIB
    nvcc = _nvcc()
    obj_dir = os.path.join(ROOT, "build", "obj" + _SUFFIX)
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in SOURCES + SELFTEST_SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, st_objs = [], []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
        (st_objs if src in SELFTEST_SOURCES else objs).append(obj)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    for cmd in ([*link, "-o", LIB, *objs],
                [*link, "-o", LIB_SELFTEST, *st_objs, "-L" + LIB_DIR, "-lnerf_b200" + _SUFFIX, "-Xlinker", "-rpath=$ORIGIN"]):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)

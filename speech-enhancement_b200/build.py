"""Build libseb200.so (sm_100a only) in-tree with nvcc.

    python speech-enhancement_b200/build.py [--force]

The objects and the shared library land next to the sources
(``speech-enhancement_b200/csrc/build/`` and ``speech-enhancement_b200/libseb200.so``); both
are git-ignored but travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libseb200.so")
OBJ_DIR = os.path.join(CSRC, "build")
SOURCES = ["core.cu", "gemm_api.cu", "gemm_train.cu", "tok_gemm.cu", "conv_y3.cu", "conv_persist.cu", "ffn_fused.cu", "dsp.cu", "norm_act.cu", "attention.cu", "attention_tc.cu", "dwconv.cu", "merge.cu", "pack.cu", "dsp_bwd.cu", "pack_dev.cu", "wgrad.cu", "train_elem.cu", "attention_train.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
NVCC_FLAGS += os.environ.get("SEB200_NVCC_EXTRA", "").split()      # experiment switches (-DSEB_...=n); part of the build digest


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def have_nvcc() -> bool:
    import shutil
    c = _nvcc()
    return os.path.exists(c) if os.path.isabs(c) else shutil.which(c) is not None


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            p = os.path.join(root, name)
            if os.path.isfile(p) and name.endswith((".cu", ".cuh", ".h")):
                h.update(name.encode())
                h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "stamp.txt")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return OUT
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", OUT, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print("built", OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)

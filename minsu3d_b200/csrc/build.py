"""Builds minsu3d_b200/lib/libb2s.so (sm_100a only) with explicit nvcc commands.

The shared library has a plain C ABI (include/b2s.h) and no torch dependency, so it is loaded
with ctypes.  It is built IN-TREE so that it travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libb2s.so")
OBJ_DIR = os.path.join(PKG, "build")
SOURCES = ["hash_map.cu", "tile_order.cu", "conv_simt.cu", "conv_tc.cu", "conv_tcp.cu", "conv_ws.cu", "wgrad_mma.cu", "wgrad_det.cu", "conv_api.cu", "fused.cu", "bn.cu", "ballquery.cu",
           "cluster.cu", "segops.cu", "postproc.cu", "augment.cu", "loss.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-diag-suppress", "177"]
# experiment switches of individual kernels (e.g. "-DB2S_TC_MIN_CTAS=4"); part of the build stamp
NVCC_FLAGS += os.environ.get("B2S_NVCC_EXTRA", "").split()


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(HERE)) + ["../../include/b2s.h"]:
        p = os.path.join(HERE, name)
        if os.path.isfile(p) and (p.endswith((".cu", ".cuh", ".h"))):
            with open(p, "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = os.path.join(LIB_DIR, "libb2s.stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)

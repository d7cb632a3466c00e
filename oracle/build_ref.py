"""Build recipe for `oracle/_ref/COMMON_OPS*.so` -- TEST INFRASTRUCTURE ONLY.

Compiles the reference's own COMMON_OPS extension from the sources where they
lie under /root/reference/minsu3d/common_ops/src (nothing is copied into this
repository) with explicit g++/nvcc commands written here (the reference's
setup.py is not run).  The three translation units are the reference's own
unity files:

    src/common_ops_api.cpp   (pybind registration, common_ops_api.cpp:6-29)
    src/common_ops.cpp       (host code,          common_ops.cpp:5-11)
    src/cuda.cu              (device code,        cuda.cu:1-8)

Outputs go only into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the .so
travels to the GPU box where /root/reference does not exist).  The module is
used by tests/ and by bench.py's cpu_baseline / --impl reference leg as the
checker and the timed baseline; the product path never imports it.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/minsu3d/common_ops/src"


def ref_so_path():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(OUT, "COMMON_OPS" + suffix)


def build(force=False, verbose=True):
    so = ref_so_path()
    stage_python(force=force, verbose=verbose)
    if os.path.exists(so) and not force:
        return so
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: only the prebuilt file is used
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + p for p in ce.include_paths("cuda")]
    inc.append("-I" + sysconfig.get_paths()["include"])
    common = ["-DTORCH_EXTENSION_NAME=COMMON_OPS", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
              "-std=c++17"]
    objs = []
    jobs = []
    for name in ("common_ops_api.cpp", "common_ops.cpp"):
        obj = os.path.join(OUT, name + ".o")
        cmd = ["g++", "-O2", "-fPIC", "-w", "-c", os.path.join(REF_SRC, name), "-o", obj] + inc + common
        jobs.append(subprocess.Popen(cmd))
        objs.append(obj)
    obj = os.path.join(OUT, "cuda.cu.o")
    cmd = ["nvcc", "-O2", "-w", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-c", os.path.join(REF_SRC, "cuda.cu"), "-o", obj] + inc + common
    jobs.append(subprocess.Popen(cmd))
    objs.append(obj)
    for j in jobs:
        if j.wait() != 0:
            raise RuntimeError("reference COMMON_OPS compile failed")
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = ["g++", "-shared", "-o", so] + objs + [
        "-L" + libdir, "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
        "-ltorch", "-ltorch_python", "-lcudart", "-Wl,-rpath," + libdir]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    if verbose:
        print("built", so)
    return so


REF_ROOT = "/root/reference"
PY_OUT = os.path.join(OUT, "refpy")


def stage_python(force=False, verbose=True):
    """Stage the reference's own Python for the hot path's callers -- minsu3d/model/**, minsu3d/common_ops/functions,
    minsu3d/loss, minsu3d/evaluation, minsu3d/util and the model / data YAML configs -- under oracle/_ref/refpy/
    (git-ignored OUTPUT directory, travels to the GPU box like the compiled reference COMMON_OPS; never committed).
    tests/test_gpu_reference_forward.py imports it from there and runs the reference's unmodified PointGroup /
    HAIS / SoftGroup `forward` + `_loss` on top of the drop-in (minsu3d_b200.install_as_reference_modules()).
    Returns the staged root or None when neither the staged copy nor /root/reference exists."""
    import shutil
    marker = os.path.join(PY_OUT, "minsu3d", "model", "pointgroup.py")
    if os.path.exists(marker) and not force:
        return PY_OUT
    src = os.path.join(REF_ROOT, "minsu3d")
    if not os.path.isdir(src):
        return None
    if os.path.isdir(PY_OUT):
        shutil.rmtree(PY_OUT)
    for sub in ("model", "common_ops/functions", "loss", "evaluation", "util", "data"):
        for root, _, files in os.walk(os.path.join(src, sub)):
            rel = os.path.relpath(root, REF_ROOT)
            for f in files:
                if f.endswith(".py"):
                    os.makedirs(os.path.join(PY_OUT, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(PY_OUT, rel, f))
    for rel in ("minsu3d/__init__.py", "minsu3d/common_ops/__init__.py"):
        dst = os.path.join(PY_OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(os.path.join(REF_ROOT, rel)):
            shutil.copyfile(os.path.join(REF_ROOT, rel), dst)
        elif not os.path.exists(dst):
            open(dst, "w").close()
    for sub in ("config/model", "config/data"):
        os.makedirs(os.path.join(PY_OUT, sub), exist_ok=True)
        for f in os.listdir(os.path.join(REF_ROOT, sub)):
            if f.endswith(".yaml"):
                shutil.copyfile(os.path.join(REF_ROOT, sub, f), os.path.join(PY_OUT, sub, f))
    if verbose:
        print("staged reference python under", PY_OUT)
    return PY_OUT


def load():
    """Import the reference COMMON_OPS module from oracle/_ref (None if absent)."""
    so = ref_so_path()
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("COMMON_OPS", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(force="--force" in sys.argv)

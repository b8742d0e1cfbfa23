"""TEST INFRASTRUCTURE ONLY — builds the reference's own native extensions as the executable oracle.

Compiles, from the sources where they lie under /root/reference (never copied into this repo):
    det3d/ops/pillar_ops/src/*.{cpp,cu}   -> oracle/_ref/pillar_cuda/pillar_cuda.so
    det3d/ops/iou3d_nms/src/*.{cpp,cu}    -> oracle/_ref/iou3d_nms_cuda/iou3d_nms_cuda.so
with the flags of the reference's setup.py files (cxx -g, nvcc -O2; det3d/ops/pillar_ops/setup.py:17-18,
det3d/ops/iou3d_nms/setup.py:13-14) for compute capability 10.0.  The reference's own build system
(setup.py build_ext) is not run; torch.utils.cpp_extension.load drives nvcc/g++ on those files directly.
Outputs go only to oracle/_ref/ (git-ignored, shipped to the GPU box).

`load_ref(name)` imports a prebuilt module from oracle/_ref without needing /root/reference.
"""
import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"

MODULES = {
    "pillar_cuda": "det3d/ops/pillar_ops/src",
    "iou3d_nms_cuda": "det3d/ops/iou3d_nms/src",
}


def _cxx_flags():
    """setup.py passes cxx ['-g']; setuptools additionally applies Python's sysconfig OPT flags (-O2/-O3),
    which matter here: without inlining, iou3d_cpu.cpp's `inline` helpers lose at link time to the
    same-named host stubs nvcc emits for the __device__ functions of iou3d_nms_kernel.cu (they exit(1))."""
    import sysconfig
    opt = (sysconfig.get_config_var("OPT") or "-O2").split()
    flags = ["-g"] + [f for f in opt if f.startswith("-O")]
    if not any(f.startswith("-O") for f in flags):
        flags.append("-O2")
    return flags


def build(name, verbose=False):
    from torch.utils.cpp_extension import load
    src_dir = os.path.join(REFERENCE, MODULES[name])
    if not os.path.isdir(src_dir):
        raise RuntimeError(f"{src_dir} not present (the reference tree only exists in the build container)")
    sources = sorted(glob.glob(os.path.join(src_dir, "*.cpp")) + glob.glob(os.path.join(src_dir, "*.cu")))
    out = os.path.join(REF_OUT, name)
    os.makedirs(out, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    return load(name=name, sources=sources, build_directory=out, extra_cflags=_cxx_flags(),
                extra_cuda_cflags=["-O2"], extra_include_paths=[src_dir], with_cuda=True,
                verbose=verbose)


def so_path(name):
    return os.path.join(REF_OUT, name, name + ".so")


def available(name):
    return os.path.exists(so_path(name))


def load_ref(name):
    """Import a prebuilt reference extension from oracle/_ref (torch must be imported first)."""
    import torch  # noqa: F401  (registers libtorch symbols the extension links against)
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, so_path(name))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def build_all(verbose=False):
    built = {}
    for name in MODULES:
        if available(name) and not os.environ.get("PN_REBUILD_REF"):
            built[name] = so_path(name)
            continue
        build(name, verbose=verbose)
        built[name] = so_path(name)
    return built


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv))

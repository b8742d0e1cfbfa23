"""TEST INFRASTRUCTURE ONLY — builds the reference's own native extensions as the executable oracle.

Compiles, from the sources where they lie under /root/reference (never copied into this repo):
    det3d/ops/pillar_ops/src/*.{cpp,cu}   -> oracle/_ref/pillar_cuda/pillar_cuda.so
    det3d/ops/iou3d_nms/src/*.{cpp,cu}    -> oracle/_ref/iou3d_nms_cuda/iou3d_nms_cuda.so
with the flags of the reference's setup.py files (cxx -g, nvcc -O2; det3d/ops/pillar_ops/setup.py:17-18,
det3d/ops/iou3d_nms/setup.py:13-14) for compute capability 10.0.  The reference's own build system
(setup.py build_ext) is not run; torch.utils.cpp_extension.load drives nvcc/g++ on those files directly.
Outputs go only to oracle/_ref/ (git-ignored, shipped to the GPU box).

`load_ref(name)` imports a prebuilt module from oracle/_ref without needing /root/reference.

`stage_python()` additionally copies the handful of reference *Python* files that drive those extensions (and the numba
CPU baseline BASELINE.json config 1 names, det3d/ops/point_cloud/point_cloud_ops.py) into oracle/_ref/py/det3d/...,
next to package __init__ stubs written here, so the GPU box can execute the reference's own Python — as the checker of
tests/test_gpu_compat.py and as bench.py's `cpu_baseline` — without /root/reference.  Like the .so files they are
git-ignored build outputs, never part of the repo's history.
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


# reference Python staged for the GPU box: relative path under /root/reference -> same path under oracle/_ref/py
PY_FILES = [
    "det3d/ops/pillar_ops/pillar_utils.py",
    "det3d/ops/pillar_ops/group_utils.py",
    "det3d/ops/pillar_ops/scatter_utils.py",
    "det3d/ops/pillar_ops/pillar_modules.py",
    "det3d/ops/iou3d_nms/__init__.py",
    "det3d/ops/iou3d_nms/iou3d_nms_utils.py",
    "det3d/ops/point_cloud/point_cloud_ops.py",
    "det3d/core/bbox/box_torch_ops.py",
    "det3d/core/utils/circle_nms_jit.py",
    "det3d/models/readers/dynamic_pillar_encoder.py",
]
# package stubs written here (the reference's own __init__ files pull in the whole framework)
PY_STUBS = {
    "det3d/__init__.py": "",
    "det3d/ops/__init__.py": "",
    "det3d/ops/pillar_ops/__init__.py": "",
    "det3d/ops/point_cloud/__init__.py": "",
    "det3d/core/__init__.py": "",
    "det3d/core/bbox/__init__.py": "",
    "det3d/core/utils/__init__.py": "",
    "det3d/models/__init__.py": "",
    "det3d/models/readers/__init__.py": "",
    # the one registry the staged reader decorates itself with (det3d/models/registry.py:3-11)
    "det3d/models/registry.py": "class _R:\n    def register_module(self, cls):\n        return cls\nREADERS = _R()\n",
    # container stand-in for spconv.pytorch.SparseConvTensor (pillar_modules.py:74)
    "spconv/__init__.py": "",
    "spconv/pytorch/__init__.py": (
        "class SparseConvTensor:\n"
        "    def __init__(self, features, indices, spatial_shape, batch_size):\n"
        "        self.features, self.indices = features, indices\n"
        "        self.spatial_shape, self.batch_size = spatial_shape, batch_size\n"),
}
PY_OUT = os.path.join(REF_OUT, "py")


def stage_python():
    import shutil
    if not os.path.isdir(os.path.join(REFERENCE, "det3d")):
        raise RuntimeError("reference tree not present")
    for rel, src in PY_STUBS.items():
        dst = os.path.join(PY_OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as fh:
            fh.write(src)
    for rel in PY_FILES:
        dst = os.path.join(PY_OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REFERENCE, rel), dst)
    return PY_OUT


def python_available():
    return all(os.path.exists(os.path.join(PY_OUT, rel)) for rel in PY_FILES)


def load_points_to_voxel():
    """the reference's numba `points_to_voxel` (det3d/ops/point_cloud/point_cloud_ops.py:112-184), imported by path
    from the staged copy"""
    path = os.path.join(PY_OUT, "det3d/ops/point_cloud/point_cloud_ops.py")
    spec = importlib.util.spec_from_file_location("pn_ref_point_cloud_ops", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.points_to_voxel


def build_all(verbose=False):
    built = {}
    for name in MODULES:
        if available(name) and not os.environ.get("PN_REBUILD_REF"):
            built[name] = so_path(name)
            continue
        build(name, verbose=verbose)
        built[name] = so_path(name)
    built["python"] = stage_python()
    return built


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv))

"""Helpers for tests/golden/make_golden.py: import the Python reference (/root/reference/det3d) in the
build container.  det3d needs `addict`, `terminaltables` and `spconv`, none installed here; tiny shim
modules are written to a temp dir and the compiled reference extensions from
oracle/_ref are registered under the names det3d imports.  Only used to *generate* fixtures."""
import os
import sys
import tempfile
import types

SHIMS = {
    "addict.py": '''
class Dict(dict):
    def __init__(self, *a, **k):
        super().__init__()
        for key, val in dict(*a, **k).items():
            self[key] = self._hook(val)
    @classmethod
    def _hook(cls, v):
        if isinstance(v, dict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._hook(x) for x in v)
        return v
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            return self.__missing__(k)
    def __missing__(self, k):
        raise KeyError(k)
    def __setattr__(self, k, v):
        self[k] = self._hook(v)
''',
    "terminaltables.py": '''
class AsciiTable:
    def __init__(self, data):
        self.table = "\\n".join(str(r) for r in data)
''',
    "spconv/__init__.py": "from . import pytorch, conv\n",
    "spconv/conv.py": "import torch.nn as nn\nclass SparseConvolution(nn.Module):\n    pass\n",
    "spconv/pytorch/__init__.py": '''
import torch.nn as nn
class SparseConvTensor:
    pass
class SparseModule(nn.Module):
    pass
class SparseSequential(nn.Sequential):
    pass
class _C(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
SubMConv2d = SubMConv3d = SparseConv2d = SparseConv3d = SparseInverseConv2d = SparseInverseConv3d = _C
class SparseReLU(nn.ReLU):
    pass
''',
}


def setup(reference="/root/reference"):
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    d = tempfile.mkdtemp(prefix="pn_shims_")
    for rel, src in SHIMS.items():
        path = os.path.join(d, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(src)
    sys.path.insert(0, d)
    sys.path.insert(0, reference)
    from oracle import build_ref
    sys.modules["det3d.ops.iou3d_nms.iou3d_nms_cuda"] = build_ref.load_ref("iou3d_nms_cuda")
    sys.modules["det3d.ops.pillar_ops.pillar_cuda"] = build_ref.load_ref("pillar_cuda")
    sys.modules["det3d.ops.roiaware_pool3d.roiaware_pool3d_cuda"] = types.ModuleType("roiaware_pool3d_cuda")

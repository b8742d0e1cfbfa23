"""pillarnet_lts_b200 — B200-native (sm_100a) PillarNet point->BEV hot path behind det3d's interfaces.

Importable as `pillarnet_lts_b200` (see the repo-root shim of that name; the directory itself is
`pillarnet-lts_b200/`).  The package holds only what the hot path needs:
  csrc/      hand-written CUDA kernels + the C ABI (include/pillarnet_b200.h)
  _lib.py    ctypes loader (no CPU fallback)
  ops.py     one Python function per C entry point
  sparse.py, layers.py, reader.py, backbone.py, neck.py, head.py, detector.py
             host-side mirror of det3d's reader / backbone / neck / head / detector interfaces
  registry.py  det3d-style registries + config loader so configs/pillarnet/*.py build unchanged
  second_stage.py, roi_head.py  Pillar R-CNN second stage, inference path (BEV fusion, RoI grid pooling, RoI head)
  synth.py   seeded synthetic LiDAR frames; dist.py frame sharding + detection gather
"""
from . import _lib  # noqa: F401
from .config import get_precision, set_precision  # noqa: F401
from . import reader, backbone, neck, head, detector, second_stage, roi_head  # noqa: F401  (fills the registries)
from .registry import (Config, build_detector, build_from_cfg, READERS, BACKBONES, NECKS, HEADS,  # noqa: F401
                       DETECTORS)

__all__ = ["set_precision", "get_precision"]

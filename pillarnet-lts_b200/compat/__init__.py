"""Signature-compatible stand-ins for the reference's two pybind extension modules on the hot path:

    pillarnet_lts_b200.compat.pillar_cuda      <->  det3d/ops/pillar_ops/pillar_cuda      (src/pillar_api.cpp:10-21)
    pillarnet_lts_b200.compat.iou3d_nms_cuda   <->  det3d/ops/iou3d_nms/iou3d_nms_cuda    (src/iou3d_nms_api.cpp:11-19)

Same function names, positional arguments (torch tensors, outputs pre-allocated by the caller and filled in place),
return values and input checks, over the C ABI of libpillarnet_b200.so — so det3d's own Python (`pillar_utils.py`,
`group_utils.py`, `scatter_utils.py`, `iou3d_nms_utils.py`, `box_torch_ops.py`) runs unmodified on this library:

    import sys
    from pillarnet_lts_b200.compat import pillar_cuda, iou3d_nms_cuda
    sys.modules["det3d.ops.pillar_ops.pillar_cuda"] = pillar_cuda
    sys.modules["det3d.ops.iou3d_nms.iou3d_nms_cuda"] = iou3d_nms_cuda

(INTEGRATION.md §C).  The fused product path (DynamicPFE / CenterHead.predict here) does not go through these.
"""
from . import iou3d_nms_cuda, pillar_cuda  # noqa: F401


def install(prefix="det3d.ops"):
    """registers the two modules under the names det3d imports (`from . import pillar_cuda` inside det3d.ops.pillar_ops
    resolves through sys.modules once the parent package is imported)"""
    import sys
    sys.modules[f"{prefix}.pillar_ops.pillar_cuda"] = pillar_cuda
    sys.modules[f"{prefix}.iou3d_nms.iou3d_nms_cuda"] = iou3d_nms_cuda

"""Multi-sweep point-cloud assembly on the device: the nuScenes branch of the reference's LoadPointCloudFromFile
(det3d/datasets/pipelines/loading.py:102-141) for clouds that are already in GPU memory (file I/O stays with the
caller).  `merge_batch` assembles a whole batch without a host sync and returns what DynamicPFE takes under
data["points_batched"]: (points (cap, 5) f32, frame_offsets (B+1,) int32 on the device)."""
import numpy as np
import torch

from . import ops


def merge_frame(key_points, sweeps, out=None, out_base=None, min_distance=1.0, n_feat=4):
    """key_points (n0, >=4) f32 CUDA; sweeps: list of dicts {"points": (n,>=4) CUDA, "transform_matrix": 4x4 or None,
    "time_lag": float} in the order they are to be appended (the reference draws the order with np.random.choice)."""
    parts = [key_points] + [s["points"] for s in sweeps]
    raw = torch.cat(parts, 0).contiguous() if len(parts) > 1 else key_points.contiguous()
    offs = np.cumsum([0] + [int(p.shape[0]) for p in parts]).tolist()
    T = [None] + [s.get("transform_matrix") for s in sweeps]
    lag = [0.0] + [float(s["time_lag"]) for s in sweeps]
    return ops.merge_sweeps(raw, offs, T, lag, min_distance=min_distance, n_feat=n_feat, out=out, out_base=out_base)


def merge_batch(frames, min_distance=1.0, n_feat=4):
    """frames: list of (key_points, sweeps).  Returns (points (cap, n_feat+1), frame_offsets (B+1,) int32 device)."""
    dev = frames[0][0].device
    cap = sum(int(k.shape[0]) + sum(int(s["points"].shape[0]) for s in sw) for k, sw in frames)
    out = torch.empty(max(cap, 1), n_feat + 1, dtype=torch.float32, device=dev)
    offsets = [torch.zeros(1, dtype=torch.int32, device=dev)]
    for key, sw in frames:
        _, total = merge_frame(key, sw, out=out, out_base=offsets[-1], min_distance=min_distance, n_feat=n_feat)
        offsets.append(total)
    return out, torch.cat(offsets)

"""Shared helpers for the -m gpu parity tests (nothing here reads /root/reference)."""
import numpy as np
import torch

from oracle import build_ref


def ref_ext(name):
    """the reference's compiled extension from oracle/_ref, or None when it did not travel"""
    if not build_ref.available(name):
        return None
    try:
        return build_ref.load_ref(name)
    except Exception:
        return None


def batch_points(frames, device="cuda"):
    """list of (Ni,5) numpy -> (points (N,5) cuda, frame_offsets (B+1,) int32 cuda)"""
    counts = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
    pts = np.concatenate(frames) if len(frames) else np.zeros((0, 5), np.float32)
    return torch.from_numpy(pts).to(device), torch.from_numpy(counts).to(device)


def rand_points(rng, n, lo=-60.0, hi=60.0):
    p = np.zeros((n, 5), np.float32)
    p[:, :2] = rng.uniform(lo, hi, (n, 2))
    p[:, 2] = rng.uniform(-3, 1, n)
    p[:, 3] = rng.random(n)
    p[:, 4] = rng.integers(0, 10, n) * 0.05
    return p


def rand_boxes(rng, n, spread=20.0, clusters=0):
    b = np.zeros((n, 7), np.float32)
    b[:, 0:2] = rng.uniform(-spread, spread, (n, 2))
    if clusters:
        c = rng.uniform(-spread, spread, (clusters, 2))
        b[:, 0:2] = c[rng.integers(0, clusters, n)] + rng.normal(0, 0.6, (n, 2))
    b[:, 2] = rng.uniform(-1, 1, n)
    b[:, 3] = rng.uniform(0.3, 5, n)
    b[:, 4] = rng.uniform(0.3, 2.5, n)
    b[:, 5] = rng.uniform(0.5, 2, n)
    b[:, 6] = rng.uniform(-4, 4, n)
    return b


def randomize_bn(model, seed=0):
    """non-trivial BN statistics so folding is exercised"""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)

"""CPU: internal consistency of the oracle (brute-force restatements, edge cases)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import pillarnet_oracle as O


def _rand_points(rng, n, lo=-60, hi=60):
    p = np.zeros((n, 5), np.float32)
    p[:, :2] = rng.uniform(lo, hi, (n, 2))
    p[:, 2] = rng.uniform(-3, 1, n)
    p[:, 3] = rng.random(n)
    p[:, 4] = rng.integers(0, 10, n) * 0.05
    return p


def test_pillarize_matches_bruteforce_dict():
    rng = np.random.default_rng(0)
    pcr, ps = [-54, -54, -5.0, 54, 54, 3.0], 0.075
    frames = [_rand_points(rng, 3000), np.zeros((0, 5), np.float32), _rand_points(rng, 1500)]
    r = O.pillarize(frames, pcr, ps)
    H, W = r["H"], r["W"]
    assert (H, W) == (1440, 1440)
    cells = {}
    for b, p in enumerate(frames):
        for q in p:
            cx = int(np.floor(np.float32(np.float32(q[0] - np.float32(pcr[0])) * (np.float32(1) / np.float32(ps)))))
            cy = int(np.floor(np.float32(np.float32(q[1] - np.float32(pcr[1])) * (np.float32(1) / np.float32(ps)))))
            if 0 <= cx < W and 0 <= cy < H:
                cells.setdefault((b, cy, cx), 0)
    want = np.array(sorted(cells), np.int32)
    assert np.array_equal(r["pillar_indices"], want)
    assert r["pts_batch_cnt"].tolist()[1] == 0
    # every point maps to the pillar holding its own cell
    pi = r["pillar_indices"][r["point_pillar_indices"]]
    assert np.array_equal(pi[:, 2], r["pts_xy"][:, 0]) and np.array_equal(pi[:, 1], r["pts_xy"][:, 1])


def test_cuda_and_true_division_modes_differ_only_on_boundaries():
    x = (np.arange(1440, dtype=np.float32) * np.float32(0.075) + np.float32(-54)).astype(np.float32)
    a = O.cell_coords(x, -54, 0.075, "cuda")
    b = O.cell_coords(x, -54, 0.075, "true_div")
    assert (a != b).sum() > 0 and np.abs(a - b).max() <= 1


def test_spatial_shapes_of_the_configs():
    assert O.bev_spatial_shape(0.075, [-54, -54, -5, 54, 54, 3]) == (1440, 1440)
    assert O.bev_spatial_shape(0.1, [-75.2, -75.2, -2, 75.2, 75.2, 4]) == (1504, 1504)
    assert O.bev_spatial_shape(0.08, [-74.88, -74.88, -2, 74.88, 74.88, 4]) == (1872, 1872)


def _random_sites(rng, B, H, W, n):
    s = set()
    while len(s) < n:
        s.add((int(rng.integers(B)), int(rng.integers(H)), int(rng.integers(W))))
    return np.array(sorted(s), np.int32)


def test_rulebooks_match_dense_equivalent_conv():
    rng = np.random.default_rng(1)
    B, H, W, C, Co = 2, 17, 14, 5, 6
    idx = _random_sites(rng, B, H, W, 120)
    feat = rng.normal(size=(len(idx), C)).astype(np.float32)
    w = rng.normal(size=(Co, 3, 3, C)).astype(np.float32)
    dense = torch.from_numpy(O.sparse_to_dense_nhwc(feat, idx, B, H, W)).permute(0, 3, 1, 2).double()
    wd = torch.from_numpy(w).permute(0, 3, 1, 2).double()
    mask = torch.zeros(B, 1, H, W, dtype=torch.float64)
    mask[idx[:, 0], 0, idx[:, 1], idx[:, 2]] = 1
    # submanifold
    nbr = O.rulebook_subm3x3(idx, H, W)
    y = O.gather_conv(feat, nbr, w.reshape(Co, 9, C))
    yd = (F.conv2d(dense, wd, padding=1) * mask).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(O.sparse_to_dense_nhwc(y, idx, B, H, W), yd, atol=1e-5)
    # strided
    oidx, nbr2, (Ho, Wo) = O.rulebook_down3x3s2(idx, H, W)
    y2 = O.gather_conv(feat, nbr2, w.reshape(Co, 9, C))
    y2d = F.conv2d(dense, wd, stride=2, padding=1)
    omask = F.max_pool2d(mask, 3, 2, 1) > 0
    assert (Ho, Wo) == tuple(y2d.shape[2:])
    got_mask = np.zeros((B, Ho, Wo), bool)
    got_mask[oidx[:, 0], oidx[:, 1], oidx[:, 2]] = True
    assert np.array_equal(got_mask, omask[:, 0].numpy())
    np.testing.assert_allclose(O.sparse_to_dense_nhwc(y2, oidx, B, Ho, Wo),
                               (y2d * omask).permute(0, 2, 3, 1).numpy(), atol=1e-5)


def test_scatter_max_zero_floor_and_empty():
    src = np.array([[-1.0, 2.0], [-3.0, 1.0], [0.5, -4.0]], np.float32)
    out = O.scatter_max(src, np.array([0, 0, 2]), 3)
    assert out.tolist() == [[0.0, 2.0], [0.0, 0.0], [0.5, 0.0]]


def test_nms_edge_cases():
    assert len(O.nms_rotated_sorted(np.zeros((0, 7), np.float32), 0.2)) == 0
    b = np.array([[0, 0, 0, 4, 2, 1.5, 0.0]] * 3, np.float32)
    assert O.nms_rotated_sorted(b, 0.2).tolist() == [0]
    b[1, 0] = 100
    assert O.nms_rotated_sorted(b, 0.2).tolist() == [0, 1]
    # known answer recorded in SURVEY §8c from the reference's CPU twin
    iou = O.boxes_iou_bev(np.array([[0, 0, 0, 4, 2, 1.5, 0]], np.float32),
                          np.array([[1, 0.5, 0, 4, 2, 1.5, 0.3]], np.float32))
    assert abs(float(iou[0, 0]) - 0.4421) < 1e-4

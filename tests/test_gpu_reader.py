"""GPU parity of dynamic pillarization + PFN/scatter-max, through the C ABI.

Bit-exact: pillar coordinates, pillar order, point->pillar indices (vs the CPU oracle, vs the same torch
expression the reference evaluates on CUDA, and vs the reference's own pillar_cuda extension from
oracle/_ref).  Tolerance: pillar features, max-abs 1e-4 (fp32; fusion only changes rounding order)."""
import numpy as np
import pytest
import torch

from oracle import pillarnet_oracle as O
from tests.gpu_util import batch_points, rand_points, ref_ext

pytestmark = pytest.mark.gpu

NUSC = dict(pcr=[-54, -54, -5.0, 54, 54, 3.0], ps=0.075)
WAYMO = dict(pcr=[-75.2, -75.2, -2, 75.2, 75.2, 4], ps=0.1)
WAYMO08 = dict(pcr=[-74.88, -74.88, -2, 74.88, 74.88, 4], ps=0.08)


def _run(frames, cfg):
    import pillarnet_lts_b200  # noqa: F401
    from pillarnet_lts_b200 import ops
    H, W = O.bev_spatial_shape(cfg["ps"], cfg["pcr"])
    pts, off = batch_points(frames)
    table, pp = ops.pillarize(pts, off, len(frames), H, W, cfg["pcr"][0], cfg["pcr"][1], cfg["ps"])
    torch.cuda.synchronize()
    return table, pp, pts, off, H, W


def _check_vs_oracle(frames, cfg):
    table, pp, pts, off, H, W = _run(frames, cfg)
    ref = O.pillarize(frames, cfg["pcr"], cfg["ps"], mode="cuda")
    n = table.count()
    assert n == len(ref["pillar_indices"])
    assert np.array_equal(table.coords[:n].cpu().numpy(), ref["pillar_indices"])
    got = pp.cpu().numpy()
    assert np.array_equal(got[got >= 0], ref["point_pillar_indices"])
    keep = np.concatenate(ref["keep"]) if frames else np.zeros(0, bool)
    assert np.array_equal(got >= 0, keep)
    return table, pp, pts, ref


def _frames(kind, n, seed0=0):
    from pillarnet_lts_b200 import synth
    return synth.make_batch(kind, n, seed0)


def test_nuscenes_frame_bit_exact():
    _check_vs_oracle(_frames("nuscenes", 1), NUSC)


def test_waymo_batch8_bit_exact():
    _check_vs_oracle(_frames("waymo", 8, 100), WAYMO)


def test_batch_with_empty_and_ragged_frames():
    rng = np.random.default_rng(5)
    frames = [rand_points(rng, 1001), np.zeros((0, 5), np.float32), rand_points(rng, 3), rand_points(rng, 777),
              np.zeros((0, 5), np.float32)]
    _check_vs_oracle(frames, NUSC)


def test_uniform_random_worst_case_rank_table():
    rng = np.random.default_rng(6)
    _check_vs_oracle([rand_points(rng, 200003)], WAYMO08)


def test_all_points_in_one_pillar_and_all_out_of_range():
    p = np.zeros((5000, 5), np.float32)
    p[:, 0] = 10.01
    p[:, 1] = -3.02
    p[:, 2] = np.linspace(-1, 1, 5000)
    t, pp, _, ref = _check_vs_oracle([p], NUSC)
    assert t.count() == 1
    q = p.copy()
    q[:, 0] = 500.0
    t, pp, _, ref = _check_vs_oracle([q], NUSC)
    assert t.count() == 0


def test_points_exactly_on_cell_and_range_borders():
    """lattice values: where true division and multiply-by-reciprocal disagree"""
    for cfg in (NUSC, WAYMO, WAYMO08):
        H, W = O.bev_spatial_shape(cfg["ps"], cfg["pcr"])
        k = np.arange(-2, W + 3, dtype=np.float64)
        xs = (cfg["pcr"][0] + k * cfg["ps"]).astype(np.float32)
        xs = np.concatenate([xs, np.nextafter(xs, np.float32(np.inf)), np.nextafter(xs, np.float32(-np.inf))])
        p = np.zeros((len(xs), 5), np.float32)
        p[:, 0] = xs
        p[:, 1] = xs[::-1]
        _check_vs_oracle([p], cfg)


def test_matches_the_torch_cuda_expression_of_the_reference():
    """dynamic_pillar_encoder.py:34-43 evaluated by torch on the GPU (not the oracle)."""
    frames = _frames("nuscenes", 2, 7)
    table, pp, pts, off, H, W = _run(frames, NUSC)
    pcr, ps = NUSC["pcr"], NUSC["ps"]
    cells = []
    for b, f in enumerate(frames):
        points = torch.from_numpy(f).cuda()
        cx = ((points[:, 0] - pcr[0]) / ps).floor().int()
        cy = ((points[:, 1] - pcr[1]) / ps).floor().int()
        m = (cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)
        cells.append(b * H * W + cy[m].long() * W + cx[m].long())
    cells = torch.cat(cells)
    uniq, inv = torch.unique(cells, sorted=True, return_inverse=True)
    n = table.count()
    assert n == uniq.numel()
    c = table.coords[:n].long()
    assert torch.equal(c[:, 0] * H * W + c[:, 1] * W + c[:, 2], uniq)
    assert torch.equal(pp[pp >= 0].long(), inv)


def _reference_group(pillar_cuda, pts_xy, pts_batch_cnt, H, W):
    """PillarQueryAndGroup.forward index part (pillar_utils.py:34-50) driven through the reference's
    own compiled kernels."""
    B = pts_batch_cnt.numel()
    n = pts_xy.shape[0]
    point_pillar_index = pts_batch_cnt.new_full((n,), -1)
    mask = pts_batch_cnt.new_zeros((B, H, W), dtype=torch.bool)
    pillar_cuda.create_point_pillar_index_stack_wrapper(pts_xy, pts_batch_cnt, mask, point_pillar_index)
    pos = torch.cumsum(mask.view(-1), dim=0, dtype=torch.int32)
    m = pos[-1].item()
    pos = pos.view(B, H, W) * mask - 1
    pillar_indices = pts_batch_cnt.new_zeros(m, 3)
    pillar_cuda.create_pillar_indices_wrapper(pos, pillar_indices)
    outs = pos.new_zeros((n,))
    pillar_cuda.gather_indice_wrapper(point_pillar_index, pos.view(-1), outs)
    return pillar_indices, outs


@pytest.mark.parametrize("kind,cfg,B", [("nuscenes", NUSC, 1), ("waymo", WAYMO, 4)])
def test_bit_exact_vs_reference_pillar_cuda(kind, cfg, B):
    pillar_cuda = ref_ext("pillar_cuda")
    if pillar_cuda is None:
        pytest.skip("oracle/_ref/pillar_cuda not built")
    frames = _frames(kind, B, 20)
    table, pp, pts, off, H, W = _run(frames, cfg)
    xy, cnt = [], []
    for f in frames:
        points = torch.from_numpy(f).cuda()
        cx = ((points[:, 0] - cfg["pcr"][0]) / cfg["ps"]).floor().int()
        cy = ((points[:, 1] - cfg["pcr"][1]) / cfg["ps"]).floor().int()
        m = (cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)
        xy.append(torch.stack((cx[m], cy[m]), dim=1))
        cnt.append(int(m.sum()))
    pts_xy = torch.cat(xy).contiguous()
    cnt = torch.tensor(cnt, dtype=torch.int32).cuda()
    ref_idx, ref_pp = _reference_group(pillar_cuda, pts_xy, cnt, H, W)
    torch.cuda.synchronize()
    n = table.count()
    assert n == ref_idx.shape[0]
    assert torch.equal(table.coords[:n], ref_idx)
    assert torch.equal(pp[pp >= 0], ref_pp)


def _pfn_params(seed=0, c=32, d=7):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(c, d, generator=g) * (2.0 / d) ** 0.5
    bn = dict(weight=torch.rand(c, generator=g) + 0.5, bias=torch.randn(c, generator=g) * 0.1,
              mean=torch.randn(c, generator=g) * 0.5, var=torch.rand(c, generator=g) + 0.5)
    return w, bn


@pytest.mark.parametrize("kind,cfg,B", [("nuscenes", NUSC, 1), ("waymo", WAYMO, 3)])
def test_pfn_scatter_max_vs_oracle_and_reference(kind, cfg, B):
    from pillarnet_lts_b200 import ops
    frames = _frames(kind, B, 40)
    table, pp, pts, ref = _check_vs_oracle(frames, cfg)
    w, bn = _pfn_params()
    inv = torch.rsqrt(bn["var"].double() + 1e-3)
    scale = (bn["weight"].double() * inv).float()
    shift = (bn["bias"].double() - bn["mean"].double() * bn["weight"].double() * inv).float()
    ps, pcr = cfg["ps"], cfg["pcr"]
    out, out_bf, arg = ops.pfn_scatter_max(pts, pp, table, pcr[0], pcr[1], ps, ps / 2.0 + pcr[0], ps / 2.0 + pcr[1],
                                           w.cuda(), scale.cuda(), shift.cuda(), want_bf16=True, want_arg=True)
    torch.cuda.synchronize()
    n = table.count()
    feat = O.point_pillar_features(ref["pts"], ref["pts_xy"], pcr, ps)
    h = O.pfn_forward(feat, w.numpy(), bn["weight"].numpy(), bn["bias"].numpy(), bn["mean"].numpy(), bn["var"].numpy())
    want = O.scatter_max(h, ref["point_pillar_indices"], n)
    got = out[:n].cpu().numpy()
    scale_ref = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= 1e-4 * scale_ref        # stated tolerance: 1e-4 rel-to-max, fp32
    assert np.abs(out_bf[:n].float().cpu().numpy() - got).max() <= 2 ** -8 * scale_ref
    # argmax really attains the max and belongs to the pillar
    a = arg[:n].cpu().numpy()
    assert (a >= 0).all()
    ppn = pp.cpu().numpy()
    assert np.array_equal(ppn[a // 32], np.repeat(np.arange(n)[:, None], 32, 1))
    # reference pipeline: torch Linear/BN/ReLU + the reference's scatter_max kernel
    pillar_cuda = ref_ext("pillar_cuda")
    if pillar_cuda is not None:
        lin = torch.nn.Linear(7, 32, bias=False).cuda()
        bnm = torch.nn.BatchNorm1d(32, momentum=0.01, eps=1e-3).cuda().eval()
        lin.weight.data.copy_(w)
        bnm.weight.data.copy_(bn["weight"]); bnm.bias.data.copy_(bn["bias"])
        bnm.running_mean.copy_(bn["mean"]); bnm.running_var.copy_(bn["var"])
        with torch.no_grad():
            hh = torch.relu(bnm(lin(torch.from_numpy(feat).cuda()))).contiguous()
        idx = torch.from_numpy(ref["point_pillar_indices"]).cuda().contiguous()
        rarg = idx.new_full((n, 32), -1)
        rout = hh.new_zeros(n, 32)
        pillar_cuda.scatter_max_wrapper(idx, hh, rarg, rout)
        torch.cuda.synchronize()
        assert (out[:n] - rout).abs().max().item() <= 1e-4 * scale_ref


@pytest.mark.parametrize("kind,cfg,B", [("nuscenes", NUSC, 2), ("waymo", WAYMO, 2)])
def test_bf16_only_fast_path_is_bit_identical_to_rounded_fp32(kind, cfg, B):
    """vector bf16 RED path (REDG.MAX.BF16x8) == round_bf16(fp32 scatter-max): rounding is monotone."""
    from pillarnet_lts_b200 import ops
    frames = _frames(kind, B, 45)
    table, pp, pts, ref = _check_vs_oracle(frames, cfg)
    w, bn = _pfn_params(3)
    scale, shift = torch.rand(32) + 0.5, torch.randn(32) * 0.1
    ps, pcr = cfg["ps"], cfg["pcr"]
    args = (pts, pp, table, pcr[0], pcr[1], ps, ps / 2.0 + pcr[0], ps / 2.0 + pcr[1], w, scale, shift)
    f32, bf_a, _ = ops.pfn_scatter_max(*args, want_bf16=True)
    none, bf_b, _ = ops.pfn_scatter_max(*args, want_bf16=True, want_f32=False)
    torch.cuda.synchronize()
    n = table.count()
    assert none is None
    assert torch.equal(bf_a[:n].view(torch.int16), f32[:n].to(torch.bfloat16).view(torch.int16))
    assert torch.equal(bf_b[:n].view(torch.int16), bf_a[:n].view(torch.int16))


def test_scatter_max_grad_routes_to_argmax():
    from pillarnet_lts_b200 import ops
    frames = _frames("nuscenes", 1, 41)
    table, pp, pts, ref = _check_vs_oracle(frames, NUSC)
    w, bn = _pfn_params(1)
    scale = torch.ones(32).cuda()
    shift = torch.zeros(32).cuda()
    ps, pcr = NUSC["ps"], NUSC["pcr"]
    out, _, arg = ops.pfn_scatter_max(pts, pp, table, pcr[0], pcr[1], ps, ps / 2.0 + pcr[0], ps / 2.0 + pcr[1],
                                      w.cuda(), scale, shift, want_arg=True)
    n = table.count()
    g = torch.randn(table.cap, 32, device="cuda")
    gs = ops.scatter_max_grad(g, arg, table, pts.shape[0])
    torch.cuda.synchronize()
    want = torch.zeros(pts.shape[0] * 32, device="cuda")
    want[arg[:n].reshape(-1).long()] = g[:n].reshape(-1)
    assert torch.equal(gs.view(-1), want)


def test_merge_sweeps_vs_reference_golden_and_feeds_pillarize(golden_dir):
    """pn_merge_sweeps vs the reference's multi-sweep loader: same points in the same order (coordinates within one
    fp32 ulp: float64 product rounded once), and a two-frame batch chains on the device into pn_pillarize."""
    import os
    from pillarnet_lts_b200 import ops, sweeps as S
    g = np.load(os.path.join(golden_dir, "sweeps.npz"))
    key = torch.from_numpy(g["raw0"]).cuda()
    sw = []
    for i in g["order"]:
        k = int(i) + 1
        T = g[f"T{k}"]
        sw.append(dict(points=torch.from_numpy(g[f"raw{k}"]).cuda(), transform_matrix=None if np.isnan(T[0, 0]) else T,
                       time_lag=float(g[f"lag{k}"])))
    out, total = S.merge_frame(key, sw)
    n = int(total.item())
    want = g["combined"]
    assert n == len(want)
    got = out[:n].cpu().numpy()
    assert np.array_equal(got[:, 3:], want[:, 3:])                      # intensity and time lag: exact
    np.testing.assert_allclose(got[:, :3], want[:, :3], rtol=1.2e-7, atol=1e-7)
    assert (got[:, :3] != want[:, :3]).mean() < 1e-3                     # and all but a few coordinates bit-equal
    # batch of two frames (the second: key frame only), offsets stay on the device
    pts, offs = S.merge_batch([(key, sw), (key[:500], [])])
    assert offs.dtype == torch.int32 and offs.cpu().tolist() == [0, n, n + 500]
    table, pp = ops.pillarize(pts, offs, 2, 1440, 1440, -54.0, -54.0, 0.075)
    assert table.count() > 1000 and int((pp[:n + 500] >= 0).sum()) > 0

"""GPU parity of rulebooks (bit-exact vs the brute-force oracle) and the gather-GEMM conv.

Tolerances: fp32 SIMT path max-abs 1e-4 relative to max|ref| (fp32 accumulate, order differs);
bf16 tcgen05 path compared with the fp32 path fed the same bf16-rounded operands: 2e-3 rel-to-max
(only accumulation order and the final bf16 rounding differ)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pillarnet_oracle as O
from tests.gpu_util import batch_points, rand_points

pytestmark = pytest.mark.gpu


def _table_from_sites(idx, B, H, W):
    """builds a RankTable on the GPU from explicit sites by pillarizing synthetic points"""
    from pillarnet_lts_b200 import ops
    ps = 1.0
    frames = []
    for b in range(B):
        s = idx[idx[:, 0] == b]
        p = np.zeros((len(s), 5), np.float32)
        p[:, 0] = s[:, 2] + 0.5
        p[:, 1] = s[:, 1] + 0.5
        frames.append(p)
    pts, off = batch_points(frames)
    table, pp = ops.pillarize(pts, off, B, H, W, 0.0, 0.0, ps)
    return table


def _random_sites(rng, B, H, W, n):
    s = set()
    while len(s) < n:
        s.add((int(rng.integers(B)), int(rng.integers(H)), int(rng.integers(W))))
    return np.array(sorted(s), np.int32)


@pytest.mark.parametrize("B,H,W,n", [(1, 16, 16, 40), (3, 37, 29, 500), (2, 128, 128, 6000), (1, 7, 5, 35)])
def test_rulebooks_bit_exact_vs_oracle(B, H, W, n):
    from pillarnet_lts_b200 import ops
    rng = np.random.default_rng(B * 1000 + n)
    idx = _random_sites(rng, B, H, W, n)
    table = _table_from_sites(idx, B, H, W)
    m = table.count()
    assert m == len(idx) and np.array_equal(table.coords[:m].cpu().numpy(), idx)
    nbr = ops.rulebook_subm3x3(table)
    assert np.array_equal(nbr[:m].cpu().numpy(), O.rulebook_subm3x3(idx, H, W))
    out_table, nbr2 = ops.rulebook_down3x3s2(table)
    oidx, onbr, (Ho, Wo) = O.rulebook_down3x3s2(idx, H, W)
    mo = out_table.count()
    assert (out_table.H, out_table.W) == (Ho, Wo) and mo == len(oidx)
    assert np.array_equal(out_table.coords[:mo].cpu().numpy(), oidx)
    assert np.array_equal(nbr2[:mo].cpu().numpy(), onbr)
    # a second level on top (stage 2 -> 3)
    t3, nbr3 = ops.rulebook_down3x3s2(out_table)
    o3, n3, _ = O.rulebook_down3x3s2(oidx, Ho, Wo)
    m3 = t3.count()
    assert np.array_equal(t3.coords[:m3].cpu().numpy(), o3) and np.array_equal(nbr3[:m3].cpu().numpy(), n3)


@pytest.mark.parametrize("B,H,W,n,levels", [(1, 16, 16, 40, 3), (3, 37, 29, 500, 3), (2, 128, 128, 6000, 3),
                                             (1, 7, 5, 35, 2), (2, 300, 212, 9000, 4), (1, 64, 64, 1, 3)])
def test_rulebook_pyramid_bit_exact_vs_oracle(B, H, W, n, levels):
    """pn_rulebook_pyramid3x3s2 (all strided levels + their submanifold tables in n_levels + 2 launches):
    coordinates, counts and both neighbour tables of every level equal the brute-force oracle's, and the
    occupancy words / prefixes equal the level-by-level entry point's."""
    from pillarnet_lts_b200 import ops
    rng = np.random.default_rng(B * 977 + n)
    idx = _random_sites(rng, B, H, W, n)
    table = _table_from_sites(idx, B, H, W)
    out = ops.rulebook_pyramid(table, levels)
    assert len(out) == levels
    prev_idx, prev_table, h, w = idx, table, H, W
    for t, nbr_down in out:
        oidx, onbr, (Ho, Wo) = O.rulebook_down3x3s2(prev_idx, h, w)
        m = t.count()
        assert (t.H, t.W, t.B) == (Ho, Wo, B) and m == len(oidx)
        assert np.array_equal(t.coords[:m].cpu().numpy(), oidx)
        assert np.array_equal(nbr_down[:m].cpu().numpy(), onbr)
        assert np.array_equal(t.subm_nbr()[:m].cpu().numpy(), O.rulebook_subm3x3(oidx, Ho, Wo))
        ref_t, ref_nbr = ops.rulebook_down3x3s2(prev_table)
        assert torch.equal(ref_t.words, t.words) and torch.equal(ref_t.prefix, t.prefix)
        assert torch.equal(ref_t.num, t.num) and torch.equal(ref_nbr[:m], nbr_down[:m])
        prev_idx, prev_table, h, w = oidx, t, Ho, Wo


def test_rulebook_pyramid_empty_and_fully_occupied_rasters():
    """edge cases of the strided levels: no active site at all (every count 0, nothing written past row 0) and a
    fully occupied raster whose width is not a multiple of 32 (every output cell active at every level; the
    word-granular occupancy kernel has to stitch runs across row boundaries inside one word)."""
    from pillarnet_lts_b200 import ops
    empty = _table_from_sites(np.zeros((0, 3), np.int32), 2, 40, 24)
    assert empty.count() == 0
    for t, nbr in ops.rulebook_pyramid(empty, 3):
        assert t.count() == 0 and int(t.words.abs().sum()) == 0
    B, H, W = 2, 33, 47
    idx = np.array([(b, y, x) for b in range(B) for y in range(H) for x in range(W)], np.int32)
    table = _table_from_sites(idx, B, H, W)
    assert table.count() == B * H * W
    prev_idx, h, w = idx, H, W
    for t, nbr_down in ops.rulebook_pyramid(table, 4):
        oidx, onbr, (Ho, Wo) = O.rulebook_down3x3s2(prev_idx, h, w)
        m = t.count()
        assert m == B * Ho * Wo == len(oidx)
        assert np.array_equal(t.coords[:m].cpu().numpy(), oidx)
        assert np.array_equal(nbr_down[:m].cpu().numpy(), onbr)
        assert np.array_equal(t.subm_nbr()[:m].cpu().numpy(), O.rulebook_subm3x3(oidx, Ho, Wo))
        prev_idx, h, w = oidx, Ho, Wo


def test_rulebook_properties_at_full_nuscenes_size():
    """size-independent properties at BASELINE size: centre tap is the identity, the table is symmetric
    (k <-> 8-k), strided outputs equal max_pool2d of the occupancy."""
    from pillarnet_lts_b200 import ops, synth
    pcr, ps = [-54, -54, -5.0, 54, 54, 3.0], 0.075
    frames = synth.make_batch("nuscenes", 2, 50)
    pts, off = batch_points(frames)
    table, pp = ops.pillarize(pts, off, 2, 1440, 1440, pcr[0], pcr[1], ps)
    m = table.count()
    nbr = ops.rulebook_subm3x3(table)[:m].long()
    ar = torch.arange(m, device="cuda")
    assert torch.equal(nbr[:, 4], ar)
    for k in range(9):
        sel = nbr[:, k] >= 0
        assert torch.equal(nbr[nbr[sel, k], 8 - k], ar[sel])
    out_table, nbr2 = ops.rulebook_down3x3s2(table)
    mo = out_table.count()
    occ = torch.zeros(2, 1, 1440, 1440, device="cuda")
    c = table.coords[:m].long()
    occ[c[:, 0], 0, c[:, 1], c[:, 2]] = 1
    pooled = F.max_pool2d(occ, 3, 2, 1)[:, 0] > 0
    got = torch.zeros_like(pooled)
    oc = out_table.coords[:mo].long()
    got[oc[:, 0], oc[:, 1], oc[:, 2]] = True
    assert torch.equal(got, pooled) and mo == int(pooled.sum())
    lin = oc[:, 0] * 720 * 720 + oc[:, 1] * 720 + oc[:, 2]
    assert bool((lin[1:] > lin[:-1]).all())  # ascending raster order
    n2 = nbr2[:mo].long()
    assert bool(((n2 >= 0).sum(1) >= 1).all()) and int(n2.max()) < m


def _conv_case(rng, B, H, W, n, cin, cout):
    idx = _random_sites(rng, B, H, W, n)
    table = _table_from_sites(idx, B, H, W)
    feat = rng.normal(size=(len(idx), cin)).astype(np.float32)
    w = (rng.normal(size=(cout, 3, 3, cin)) * 0.2).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.normal(0, 0.2, cout).astype(np.float32)
    return idx, table, feat, w, scale, shift


@pytest.mark.parametrize("cin,cout", [(32, 32), (32, 64), (64, 64), (5, 7)])
def test_sparse_conv_simt_vs_oracle(cin, cout):
    from pillarnet_lts_b200 import ops
    rng = np.random.default_rng(cin * 100 + cout)
    B, H, W = 2, 24, 20
    idx, table, feat, w, scale, shift = _conv_case(rng, B, H, W, 300, cin, cout)
    m = table.count()
    x = torch.zeros(table.cap, cin, device="cuda")
    x[:m] = torch.from_numpy(feat).cuda()
    res = torch.randn(table.cap, cout, device="cuda")
    wt = torch.from_numpy(w.reshape(cout, -1)).cuda()
    out = torch.full((table.cap, cout), 7.0, device="cuda")
    ops.conv_gather(x, wt, table.subm_nbr(), 9, cin, cout, out, scale=torch.from_numpy(scale).cuda(),
                    shift=torch.from_numpy(shift).cuda(), residual=res, relu=True, num=table.num)
    want = O.gather_conv(feat, O.rulebook_subm3x3(idx, H, W), w.reshape(cout, 9, cin), scale, shift,
                         res[:m].cpu().numpy(), relu=True)
    got = out[:m].cpu().numpy()
    assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max())
    assert bool((out[m:] == 7.0).all())  # rows past the device count are untouched
    # strided
    ot, nbr2 = ops.rulebook_down3x3s2(table)
    mo = ot.count()
    out2 = torch.empty(ot.cap, cout, device="cuda")
    ops.conv_gather(x, wt, nbr2, 9, cin, cout, out2, num=ot.num)
    oidx, onbr, _ = O.rulebook_down3x3s2(idx, H, W)
    want2 = O.gather_conv(feat, onbr, w.reshape(cout, 9, cin))
    assert np.abs(out2[:mo].cpu().numpy() - want2).max() <= 1e-4 * max(1.0, np.abs(want2).max())


def test_dense_conv_tables_vs_torch():
    """3x3 s1/s2 and ConvTranspose2d(2,2) through the static gather tables vs torch (fp32, no TF32)."""
    from pillarnet_lts_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W, cin, cout = 2, 13, 10, 24, 40
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    rows = x.permute(0, 2, 3, 1).contiguous().view(B * H * W, cin)
    for stride in (1, 2):
        conv = torch.nn.Conv2d(cin, cout, 3, stride, 1, bias=False).cuda()
        w2d = conv.weight.detach().permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
        nbr = ops.dense_nbr_table(0, B, H, W, stride, x.device)
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        out = torch.empty(B * Ho * Wo, cout, device="cuda")
        ops.conv_gather(rows, w2d, nbr, 9, cin, cout, out)
        want = conv(x).permute(0, 2, 3, 1).reshape(B * Ho * Wo, cout)
        assert (out - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())
    de = torch.nn.ConvTranspose2d(cin, cout, 2, 2, bias=False).cuda()
    w2d = de.weight.detach().permute(1, 2, 3, 0).reshape(cout, -1).contiguous()
    nbr = ops.dense_nbr_table(1, B, H, W, 2, x.device)
    out = torch.empty(B * 4 * H * W, cout, device="cuda")
    ops.conv_gather(rows, w2d, nbr, 4, cin, cout, out)
    want = de(x).permute(0, 2, 3, 1).reshape(-1, cout)
    assert (out - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())


def test_sparse_to_dense_and_channel_offset():
    from pillarnet_lts_b200 import ops
    rng = np.random.default_rng(9)
    B, H, W, C = 2, 11, 9, 8
    idx = _random_sites(rng, B, H, W, 60)
    table = _table_from_sites(idx, B, H, W)
    m = table.count()
    for dt in (torch.float32, torch.bfloat16):
        feat = torch.zeros(table.cap, C, device="cuda", dtype=dt)
        feat[:m] = torch.randn(m, C, device="cuda").to(dt)
        wide = torch.full((B * H * W, 2 * C), 3.0, device="cuda", dtype=dt)
        ops.sparse_to_dense(feat, table, C, out=wide, out_coff=C)
        want = torch.zeros(B, H, W, C, device="cuda", dtype=dt)
        want[idx[:, 0], idx[:, 1], idx[:, 2]] = feat[:m]
        assert torch.equal(wide[:, C:], want.view(-1, C)) and bool((wide[:, :C] == 3.0).all())


@pytest.mark.parametrize("cin,cout,taps_case", [(32, 32, "subm"), (64, 128, "down"), (256, 256, "dense"),
                                                (128, 64, "dense"), (64, 3, "dense"), (512, 256, "dense"),
                                                (256, 128, "deconv")])
def test_tcgen05_conv_vs_simt_same_operands(cin, cout, taps_case):
    """bf16 tensor-core path vs the fp32-FMA path on identical bf16 operands."""
    from pillarnet_lts_b200 import ops
    from pillarnet_lts_b200._lib import PN_IMPL_SIMT, PN_IMPL_TCGEN05
    rng = np.random.default_rng(cin + cout)
    if taps_case in ("subm", "down"):
        B, H, W = 2, 96, 96
        idx = _random_sites(rng, B, H, W, 3000)
        table = _table_from_sites(idx, B, H, W)
        if taps_case == "subm":
            nbr, num, cap, taps = table.subm_nbr(), table.num, table.cap, 9
        else:
            ot, nbr = ops.rulebook_down3x3s2(table)
            num, cap, taps = ot.num, ot.cap, 9
        rows_in = table.cap
    elif taps_case == "dense":
        B, H, W = 2, 20, 24
        nbr, num, cap, taps, rows_in = ops.dense_nbr_table(0, B, H, W, 1, "cuda"), None, B * H * W, 9, B * H * W
    else:
        B, H, W = 2, 10, 12
        nbr, num, cap, taps, rows_in = ops.dense_nbr_table(1, B, H, W, 2, "cuda"), None, B * 4 * H * W, 4, B * H * W
    x = torch.randn(rows_in, cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, taps * cin, device="cuda") * (1.0 / (taps * cin) ** 0.5))
    wp = ops.pack_weight_bf16(w.contiguous())
    scale = (torch.rand(cout, device="cuda") + 0.5)
    shift = torch.randn(cout, device="cuda") * 0.1
    res = torch.randn(cap, cout, device="cuda").to(torch.bfloat16)
    outs = []
    for impl in (PN_IMPL_SIMT, PN_IMPL_TCGEN05):
        out = torch.zeros(cap, cout, device="cuda", dtype=torch.bfloat16)
        ops.conv_gather(x, wp, nbr, taps, cin, cout, out, scale=scale, shift=shift, residual=res, relu=True,
                        num=num, rows_cap=cap, impl=impl)
        torch.cuda.synchronize()
        outs.append(out.float())
    ref = outs[0]
    # both paths round an (almost) identical fp32 value to bf16: allow one bf16 ulp (2^-7 relative)
    assert (outs[1] - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())
    # f32 output variant (head maps)
    o32 = [torch.zeros(cap, cout, device="cuda") for _ in range(2)]
    for o, impl in zip(o32, (PN_IMPL_SIMT, PN_IMPL_TCGEN05)):
        ops.conv_gather(x, wp, nbr, taps, cin, cout, o, scale=scale, shift=shift, num=num, rows_cap=cap, impl=impl)
    torch.cuda.synchronize()
    assert (o32[1] - o32[0]).abs().max().item() <= 2e-3 * max(1.0, o32[0].abs().max().item())


def test_grouped_small_cout_conv_vs_torch():
    """all CenterHead final convs in one launch (fp32 accumulation on bf16 inputs) vs F.conv2d, tol 1e-4."""
    from pillarnet_lts_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W, hc = 2, 45, 37, 64
    couts = [2, 1, 3, 2, 2, 1, 4]
    G = len(couts)
    x = torch.randn(B * H * W, G * hc, device="cuda", generator=g).to(torch.bfloat16)
    ws = [torch.randn(c, hc, 3, 3, device="cuda", generator=g) * 0.1 for c in couts]
    bs = [torch.randn(c, device="cuda", generator=g) for c in couts]
    desc, chunks, off, col = [], [], 0, 0
    for i, (w, b) in enumerate(zip(ws, bs)):
        wf = w.permute(0, 2, 3, 1).reshape(-1)
        desc.append([i * hc, w.shape[0], off, off + wf.numel(), col])
        chunks += [wf, b]
        off += wf.numel() + b.numel()
        col += w.shape[0]
    groups = torch.tensor(desc, dtype=torch.int32).cuda()
    wbuf = torch.cat(chunks).contiguous()
    out = torch.full((B * H * W, col + 3), -9.0, device="cuda")
    ops.conv3x3_small_cout(x, x.stride(0), hc, B, H, W, groups, G, wbuf, out)
    torch.cuda.synchronize()
    xin = x.float().view(B, H, W, G * hc).permute(0, 3, 1, 2)
    c0 = 0
    for i, (w, b) in enumerate(zip(ws, bs)):
        want = torch.nn.functional.conv2d(xin[:, i * hc:(i + 1) * hc], w, b, padding=1).permute(0, 2, 3, 1)
        got = out[:, c0:c0 + w.shape[0]].view(B, H, W, -1)
        assert (got - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())
        c0 += w.shape[0]
    assert bool((out[:, col:] == -9.0).all())


# ---- window-staged submanifold kernel (conv_win_tc.cu) ---------------------------------------------------------------
def _clustered_sites(rng, B, H, W, n):
    """LiDAR-like occupancy: arcs + blobs, so raster neighbours exist (random sites have almost none)"""
    ys, xs, bs = [], [], []
    for b in range(B):
        for r in rng.uniform(5, min(H, W) / 2 - 2, 14):
            th = rng.uniform(0, 2 * np.pi, max(8, int(n / 14 / B)))
            ys.append(np.clip(np.round(H / 2 + r * np.sin(th)), 0, H - 1))
            xs.append(np.clip(np.round(W / 2 + r * np.cos(th)), 0, W - 1))
            bs.append(np.full(len(th), b))
        cy, cx = rng.integers(8, H - 8, 10), rng.integers(8, W - 8, 10)
        for y0, x0 in zip(cy, cx):
            yy, xx = np.meshgrid(np.arange(-5, 6), np.arange(-7, 8), indexing="ij")
            keep = rng.random(yy.shape) < 0.8
            ys.append((y0 + yy)[keep]); xs.append(np.clip((x0 + xx)[keep], 0, W - 1)); bs.append(np.full(int(keep.sum()), b))
    idx = np.unique(np.stack([np.concatenate(bs), np.concatenate(ys), np.concatenate(xs)], 1).astype(np.int64), axis=0)
    return idx


@pytest.mark.parametrize("c,HW,B", [(32, 200, 2), (64, 160, 1), (128, 120, 2), (64, 40, 1), (256, 96, 2)])
def test_window_staged_subm_conv_equals_gather_kernel(c, HW, B):
    """conv_win_tc.cu (nbr_kind = PN_NBR_SUBM_SORTED) against the gather kernel (nbr_kind = 0) and the fp32-FMA path
    on identical bf16 operands, with residual + ReLU, partial last tiles and rows beyond the live count untouched."""
    from pillarnet_lts_b200 import ops
    from pillarnet_lts_b200._lib import PN_IMPL_SIMT, PN_IMPL_TCGEN05, PN_NBR_SUBM_SORTED
    rng = np.random.default_rng(c + HW)
    idx = _clustered_sites(rng, B, HW, HW, 6000)
    table = _table_from_sites(idx, B, HW, HW)
    m = table.count()
    nbr = table.subm_nbr()
    cap = table.cap
    x = torch.randn(cap, c, device="cuda").to(torch.bfloat16)
    w = torch.randn(c, 9 * c, device="cuda") * (1.0 / (9 * c) ** 0.5)
    wp = ops.pack_weight_bf16(w.contiguous())
    scale, shift = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    res = torch.randn(cap, c, device="cuda").to(torch.bfloat16)
    plan = table.subm_plan()
    outs = {}
    for name, impl, kind in (("simt", PN_IMPL_SIMT, 0), ("gather", PN_IMPL_TCGEN05, 0),
                             ("window", PN_IMPL_TCGEN05, PN_NBR_SUBM_SORTED)):
        out = torch.full((cap, c), 7.0, device="cuda", dtype=torch.bfloat16)
        ops.conv_gather(x, wp, nbr, 9, c, c, out, scale=scale, shift=shift, residual=res, relu=True, num=table.num,
                        rows_cap=cap, impl=impl, nbr_kind=kind, nbr_plan=plan if kind else None)
        torch.cuda.synchronize()
        outs[name] = out.float()
    ref = outs["simt"]
    assert m > 500 and bool((outs["window"][m:] == 7.0).all())          # rows past the live count are not written
    tol = 1e-2 * max(1.0, ref.abs().max().item())                       # one bf16 ulp of the largest value
    assert (outs["window"][:m] - ref[:m]).abs().max().item() <= tol
    assert (outs["window"][:m] - outs["gather"][:m]).abs().max().item() <= tol
    # same fp32 accumulation order when there is one K chunk per tap: bit-equal to the gather kernel
    if c <= 64:
        assert torch.equal(outs["window"][:m], outs["gather"][:m])
    # no residual / no ReLU / no affine
    o1 = torch.zeros(cap, c, device="cuda", dtype=torch.bfloat16)
    o2 = torch.zeros(cap, c, device="cuda", dtype=torch.bfloat16)
    ops.conv_gather(x, wp, nbr, 9, c, c, o1, num=table.num, rows_cap=cap, impl=PN_IMPL_TCGEN05, nbr_kind=0)
    ops.conv_gather(x, wp, nbr, 9, c, c, o2, num=table.num, rows_cap=cap, impl=PN_IMPL_TCGEN05,
                    nbr_kind=PN_NBR_SUBM_SORTED, nbr_plan=plan)
    torch.cuda.synchronize()
    assert (o1.float() - o2.float())[:m].abs().max().item() <= 1e-2 * max(1.0, o1.float().abs().max().item())


def test_window_staged_kernel_is_correct_for_any_rulebook():
    """the hint never changes results: a scrambled (non-raster) rulebook sends every neighbour through the
    out-of-window global fetch of conv_win_tc.cu; an empty site set and a one-row set run too."""
    from pillarnet_lts_b200 import ops
    from pillarnet_lts_b200._lib import PN_IMPL_TCGEN05, PN_NBR_SUBM_SORTED
    g = torch.Generator(device="cuda").manual_seed(11)
    rows, c = 3000, 64
    nbr = torch.randint(-1, rows, (rows, 9), device="cuda", generator=g, dtype=torch.int32)
    nbr[torch.rand(rows, 9, device="cuda", generator=g) < 0.3] = -1
    x = torch.randn(rows, c, device="cuda", generator=g).to(torch.bfloat16)
    wp = ops.pack_weight_bf16((torch.randn(c, 9 * c, device="cuda", generator=g) / 24.0).contiguous())
    outs = []
    plan = ops.conv_window_plan(nbr, None, rows)
    for kind in (0, PN_NBR_SUBM_SORTED):
        out = torch.zeros(rows, c, device="cuda", dtype=torch.bfloat16)
        ops.conv_gather(x, wp, nbr, 9, c, c, out, relu=True, impl=PN_IMPL_TCGEN05, nbr_kind=kind,
                        nbr_plan=plan if kind else None)
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    for n_live in (0, 1, 65):
        num = torch.tensor([n_live], dtype=torch.int32, device="cuda")
        out = torch.full((rows, c), 3.0, device="cuda", dtype=torch.bfloat16)
        ref = torch.full((rows, c), 3.0, device="cuda", dtype=torch.bfloat16)
        nb = nbr.clone()
        nb[nb >= max(n_live, 1)] = -1
        ops.conv_gather(x, wp, nb, 9, c, c, out, num=num, rows_cap=rows, impl=PN_IMPL_TCGEN05,
                        nbr_kind=PN_NBR_SUBM_SORTED, nbr_plan=ops.conv_window_plan(nb, num, rows))
        ops.conv_gather(x, wp, nb, 9, c, c, ref, num=num, rows_cap=rows, impl=PN_IMPL_TCGEN05, nbr_kind=0)
        torch.cuda.synchronize()
        assert torch.equal(out, ref)

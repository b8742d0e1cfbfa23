"""GPU parity of the padded-layout dense 3x3 tensor-core conv (pn_conv_dense3x3) vs torch F.conv2d on the
same bf16-rounded operands (fp32 accumulate both sides): 2e-3 rel-to-max for f32 output, one bf16 ulp
(1e-2 rel-to-max) for bf16 output."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _pad_rows(x_nhwc):
    """(B,H,W,C) -> zero-padded rows (B*(H+2)*(W+2), C)"""
    B, H, W, C = x_nhwc.shape
    p = torch.zeros(B, H + 2, W + 2, C, dtype=x_nhwc.dtype, device=x_nhwc.device)
    p[:, 1:-1, 1:-1] = x_nhwc
    return p.view(-1, C).contiguous()


def _run(B, H, W, cin, cout, tile_hint, compact, out_dtype, in_extra=0, relu=True):
    from pillarnet_lts_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + cin + cout)
    x = torch.randn(B, H, W, cin + in_extra, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (1.0 / (9 * cin) ** 0.5)
    wp = ops.pack_weight_bf16(w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous())
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(cout, device="cuda", generator=g) * 0.1
    rows = _pad_rows(x)
    n_out = B * H * W if compact else rows.shape[0]
    out = torch.full((n_out, cout + 8), 5.0, device="cuda", dtype=out_dtype)
    ops.conv_dense3x3(rows, in_extra, cin, B, H, W, wp, cout, out, scale=scale, shift=shift, out_coff=8,
                      out_compact=compact, relu=relu, tile_hint=tile_hint)
    torch.cuda.synchronize()
    xin = x[..., in_extra:].float().permute(0, 3, 1, 2)
    want = F.conv2d(xin, wp[:, :9 * cin].float().view(cout, 3, 3, cin).permute(0, 3, 1, 2), padding=1)
    want = want * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if relu:
        want = want.relu()
    want = want.permute(0, 2, 3, 1)
    if compact:
        got = out[:, 8:].float().view(B, H, W, cout)
    else:
        full = out[:, 8:].float().view(B, H + 2, W + 2, cout)
        assert float(full[:, 0].abs().max()) == 0 and float(full[:, -1].abs().max()) == 0
        assert float(full[:, :, 0].abs().max()) == 0 and float(full[:, :, -1].abs().max()) == 0
        got = full[:, 1:-1, 1:-1]
    assert bool((out[:, :8] == 5.0).all())
    err = (got - want).abs().max().item() / max(1.0, want.abs().max().item())
    return err


def test_descriptor_window_mode_probe():
    """which UMMA descriptor encoding addresses a row-shifted window of a 128B-swizzled tile"""
    errs = {}
    for mode in (0, 0x100):
        errs[mode] = _run(1, 20, 24, 64, 64, 4 | mode, False, torch.float32)
    print("dense conv window probe: base_offset=0 err %.3g, base_offset=(addr>>7)&7 err %.3g" % (errs[0], errs[0x100]))
    assert min(errs.values()) <= 2e-3


@pytest.mark.parametrize("cluster", [0x800, 0x200, 0x400, 0x1000])   # no cluster / multicast over 2, 4 CTAs / cta_group::2 pair
@pytest.mark.parametrize("tile", [1, 2, 3, 4])
@pytest.mark.parametrize("B,H,W,cin,cout", [(1, 20, 24, 64, 64), (2, 33, 17, 128, 256), (1, 45, 45, 256, 160),
                                            (3, 61, 50, 64, 320)])
def test_dense_conv_vs_torch(cluster, tile, B, H, W, cin, cout):
    if tile in (1, 2) and cout <= 128:
        pytest.skip("BN=256 tiles are not offered for cout <= 128")
    assert _run(B, H, W, cin, cout, tile | cluster, False, torch.float32) <= 2e-3
    assert _run(B, H, W, cin, cout, tile | cluster, True, torch.bfloat16, in_extra=64) <= 1e-2


def test_dense_conv_auto_tile_full_size():
    assert _run(1, 180, 180, 256, 256, 0, False, torch.bfloat16) <= 1e-2
    assert _run(1, 90, 90, 256, 256, 0, False, torch.bfloat16) <= 1e-2
    assert _run(1, 180, 180, 64, 2304, 0, True, torch.bfloat16, relu=True) <= 1e-2


def test_grouped_small_cout_dense_conv_vs_torch():
    """all final head convs in one tensor-core launch vs F.conv2d on the same bf16 operands (2e-3 rel-to-max)."""
    from pillarnet_lts_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(7)
    B, H, W, hc = 2, 37, 29, 64
    couts = [2, 1, 3, 2, 2, 1, 4, 16]
    G = len(couts)
    x = torch.randn(B, H, W, G * hc, device="cuda", generator=g).to(torch.bfloat16)
    rows = _pad_rows(x)
    ws = [torch.randn(c, hc, 3, 3, device="cuda", generator=g) * 0.1 for c in couts]
    bs = [torch.randn(c, device="cuda", generator=g) for c in couts]
    wf = torch.zeros(G * 16, 9 * hc, device="cuda")
    bf = torch.zeros(G * 16, device="cuda")
    tab, col = [], 3
    for i, (w, b) in enumerate(zip(ws, bs)):
        wf[i * 16:i * 16 + w.shape[0]] = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
        bf[i * 16:i * 16 + w.shape[0]] = b
        tab.append([col, w.shape[0]])
        col += w.shape[0]
    wg = ops.pack_weight_bf16(wf)
    tabd = torch.tensor(tab, dtype=torch.int32).cuda()
    out = torch.full((B * H * W, col + 2), -9.0, device="cuda")
    ops.conv_dense3x3_grouped(rows, 0, hc, G, B, H, W, wg, bf, tabd, out, out_compact=True)
    torch.cuda.synchronize()
    xin = x.float().permute(0, 3, 1, 2)
    c0 = 3
    for i, (w, b) in enumerate(zip(ws, bs)):
        wq = wg[i * 16:i * 16 + w.shape[0], :9 * hc].float().view(w.shape[0], 3, 3, hc).permute(0, 3, 1, 2)
        want = F.conv2d(xin[:, i * hc:(i + 1) * hc], wq, b, padding=1).permute(0, 2, 3, 1)
        got = out[:, c0:c0 + w.shape[0]].view(B, H, W, -1)
        assert (got - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item()), i
        c0 += w.shape[0]
    assert bool((out[:, :3] == -9.0).all()) and bool((out[:, col:] == -9.0).all())


def test_planar_intermediate_layout_matches_interleaved():
    """pn_conv_dense3x3(out_group_cols=64) writes one contiguous padded map per 64 output channels, and
    pn_conv_dense3x3_grouped(in_planar=1) reads it: both must equal the interleaved-layout results bit for bit."""
    from pillarnet_lts_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    B, H, W, cin, G, hc = 2, 23, 31, 64, 5, 64
    x = torch.randn(B, H, W, cin, device="cuda", generator=g).to(torch.bfloat16)
    rows = _pad_rows(x)
    n_rows = rows.shape[0]
    w1 = ops.pack_weight_bf16(torch.randn(G * hc, 9 * cin, device="cuda", generator=g) * 0.05)
    sc = torch.rand(G * hc, device="cuda", generator=g) + 0.5
    sh = torch.randn(G * hc, device="cuda", generator=g) * 0.1
    inter = torch.empty(n_rows, G * hc, device="cuda", dtype=torch.bfloat16)
    ops.conv_dense3x3(rows, 0, cin, B, H, W, w1, G * hc, inter, scale=sc, shift=sh, relu=True)
    planar = torch.full((G * n_rows, hc), 3.0, device="cuda", dtype=torch.bfloat16)
    ops.conv_dense3x3(rows, 0, cin, B, H, W, w1, G * hc, planar, scale=sc, shift=sh, relu=True, out_group_cols=hc)
    torch.cuda.synchronize()
    assert torch.equal(planar.view(G, n_rows, hc).permute(1, 0, 2).reshape(n_rows, G * hc), inter)
    couts = [2, 1, 3, 2, 4]
    wf = torch.zeros(G * 16, 9 * hc, device="cuda")
    bf = torch.zeros(G * 16, device="cuda")
    tab, col = [], 0
    for i, c in enumerate(couts):
        wf[i * 16:i * 16 + c] = torch.randn(c, 9 * hc, device="cuda", generator=g) * 0.1
        bf[i * 16:i * 16 + c] = torch.randn(c, device="cuda", generator=g)
        tab.append([col, c])
        col += c
    wg = ops.pack_weight_bf16(wf)
    tabd = torch.tensor(tab, dtype=torch.int32).cuda()
    out_a = torch.full((B * H * W, col), -9.0, device="cuda")
    out_b = torch.full((B * H * W, col), -9.0, device="cuda")
    ops.conv_dense3x3_grouped(inter, 0, hc, G, B, H, W, wg, bf, tabd, out_a, out_compact=True)
    ops.conv_dense3x3_grouped(planar, 0, hc, G, B, H, W, wg, bf, tabd, out_b, out_compact=True, in_planar=True)
    torch.cuda.synchronize()
    assert torch.equal(out_a, out_b)
    assert float(out_a.abs().max()) > 0.1


@pytest.mark.parametrize("B,H,W,G", [(2, 23, 31, 5), (1, 180, 180, 36), (3, 7, 5, 3), (1, 40, 253, 2), (2, 64, 64, 150)])
def test_grouped_last_conv_as_gemm_plus_shifted_sums(B, H, W, G):
    """pn_conv_dense3x3_grouped_shift (1x1 GEMM per branch + nine shifted sums) on the planar padded intermediate vs
    F.conv2d on the same bf16 operands (2e-3 rel-to-max: fp32 accumulation, different summation order) and vs the
    implicit-GEMM grouped kernel; strips with halos, several frames, more branches than SMs, columns outside the
    groups untouched; maps wider than the halo are refused."""
    from pillarnet_lts_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(B * 100 + H + G)
    hc = 64
    x = (torch.randn(G, B, H, W, hc, device="cuda", generator=g)).relu().to(torch.bfloat16)
    n_pos = B * (H + 2) * (W + 2)
    planar = torch.stack([_pad_rows(x[i]) for i in range(G)]).reshape(G * n_pos, hc).contiguous()
    couts = [(1, 2, 3)[i % 3] for i in range(G)]
    ws = [torch.randn(c, hc, 3, 3, device="cuda", generator=g) * 0.1 for c in couts]
    bs = [torch.randn(c, device="cuda", generator=g) for c in couts]
    w32 = torch.zeros(G, 32, hc, device="cuda")
    b4 = torch.zeros(G, 4, device="cuda")
    wf = torch.zeros(G * 16, 9 * hc, device="cuda")
    bf = torch.zeros(G * 16, device="cuda")
    tab, col = [], 2
    for i, (w, b) in enumerate(zip(ws, bs)):
        c = w.shape[0]
        w32[i, :27] = torch.nn.functional.pad(w.permute(2, 3, 0, 1).reshape(9, c, hc), (0, 0, 0, 3 - c)).reshape(27, hc)
        b4[i, :c] = b
        wf[i * 16:i * 16 + c] = w.permute(0, 2, 3, 1).reshape(c, -1)
        bf[i * 16:i * 16 + c] = b
        tab.append([col, c])
        col += c
    wsh = w32.reshape(G * 32, hc).to(torch.bfloat16).contiguous()
    tabd = torch.tensor(tab, dtype=torch.int32).cuda()
    out = torch.full((B * H * W, col + 3), -9.0, device="cuda")
    ops.conv_dense3x3_grouped_shift(planar, G, B, H, W, wsh, b4.reshape(-1).contiguous(), tabd, out)
    ref = torch.full((B * H * W, col + 3), -9.0, device="cuda")
    ops.conv_dense3x3_grouped(planar, 0, hc, G, B, H, W, ops.pack_weight_bf16(wf), bf, tabd, ref, out_compact=True,
                              in_planar=True)
    torch.cuda.synchronize()
    assert bool((out[:, :2] == -9.0).all()) and bool((out[:, col:] == -9.0).all())
    c0 = 2
    for i, (w, b) in enumerate(zip(ws, bs)):
        c = w.shape[0]
        wq = w.to(torch.bfloat16).float()
        want = F.conv2d(x[i].float().permute(0, 3, 1, 2), wq, b, padding=1).permute(0, 2, 3, 1)
        got = out[:, c0:c0 + c].view(B, H, W, c)
        tol = 2e-3 * max(1.0, want.abs().max().item())
        assert (got - want).abs().max().item() <= tol, i
        assert (got - ref[:, c0:c0 + c].view(B, H, W, c)).abs().max().item() <= tol, i
        c0 += c


def test_grouped_shift_refuses_maps_wider_than_its_halo():
    from pillarnet_lts_b200 import ops
    G, B, H, W, hc = 1, 1, 4, 300, 64
    planar = torch.zeros(G * B * (H + 2) * (W + 2), hc, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(NotImplementedError):
        ops.conv_dense3x3_grouped_shift(planar, G, B, H, W, torch.zeros(32, hc, device="cuda", dtype=torch.bfloat16),
                                        torch.zeros(4, device="cuda"), torch.tensor([[0, 1]], dtype=torch.int32).cuda(),
                                        torch.zeros(B * H * W, 1, device="cuda"))


@pytest.mark.parametrize("B,H,W,cin,cout,coff,width", [(1, 12, 9, 64, 64, 0, 64), (2, 45, 45, 256, 256, 256, 512),
                                                        (1, 90, 90, 256, 128, 8, 136), (1, 5, 7, 128, 48, 0, 48)])
def test_deconv2x2_gemm_form_vs_torch_and_gather_form(B, H, W, cin, cout, coff, width):
    """ConvTranspose2d(k=2,s=2)+BN+ReLU (necks/rpn.py:150-154) as one GEMM with a scattering epilogue
    (pn_conv_args.deconv_*): equals torch on the same bf16-rounded operands within one bf16 ulp (1e-2 rel-to-max),
    equals the 4-tap gather formulation bit for bit (same products, same fp32 accumulation order over Cin),
    keeps the zero border of the padded output map and leaves the neighbouring columns untouched."""
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import config
    from pillarnet_lts_b200.layers import DenseMap, dense_deconv2x2
    P.set_precision("bf16")
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(H * 100 + cin + cout)
    conv = torch.nn.ConvTranspose2d(cin, cout, 2, stride=2, bias=False).cuda()
    bn = torch.nn.BatchNorm2d(cout, eps=1e-3).cuda().eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.1)
        bn.running_var.uniform_(0.5, 1.5)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.1)
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    xm = DenseMap(_pad_rows(x), B, H, W, cin, 0, 1)
    outs = []
    for gemm in (True, False):
        config.set_deconv_gemm(gemm)
        rows = torch.full((B * (2 * H + 2) * (2 * W + 2), width), 3.0, device="cuda", dtype=torch.bfloat16)
        try:
            dense_deconv2x2(xm, conv, bn, relu=True, out=rows, out_coff=coff)
        finally:
            config.set_deconv_gemm(True)
        torch.cuda.synchronize()
        outs.append(rows)
    got, ref_gather = outs
    assert torch.equal(got, ref_gather)
    full = got.view(B, 2 * H + 2, 2 * W + 2, width)
    if coff:
        assert bool((full[..., :coff] == 3.0).all())
    assert bool((full[..., coff + cout:] == 3.0).all())
    y = full[..., coff:coff + cout].float()
    assert float(y[:, 0].abs().max()) == 0 and float(y[:, -1].abs().max()) == 0
    assert float(y[:, :, 0].abs().max()) == 0 and float(y[:, :, -1].abs().max()) == 0
    with torch.no_grad():
        w = conv.weight.to(torch.bfloat16).float()
        want = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w, stride=2)
        want = bn(want).relu().permute(0, 2, 3, 1)
    err = (y[:, 1:-1, 1:-1] - want).abs().max().item() / max(1.0, want.abs().max().item())
    assert err < 1e-2, err

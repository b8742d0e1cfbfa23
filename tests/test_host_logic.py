"""CPU tests of host-side logic added around the CUDA path (no device needed)."""
import torch
import torch.nn.functional as F


class _FakeTable:
    def __init__(self, B, H, W, cap, n):
        self.B, self.H, self.W, self.cap, self._n = B, H, W, cap, n

    def count(self):
        return self._n


def test_observed_row_counts_steer_the_tile_hint():
    """backbone._rows_hint: 25 % of the cells until the engine has measured a batch, the observed count afterwards
    (clamped to the capacity); empty observations are ignored."""
    from pillarnet_lts_b200 import backbone
    saved = dict(backbone._observed_rows)
    backbone._observed_rows.clear()
    try:
        t = _FakeTable(1, 180, 180, 32400, 10283)
        assert backbone._rows_hint(t) == 180 * 180 // 4
        backbone.observe_rows_begin()
        backbone._observing.extend([t, _FakeTable(1, 90, 90, 8100, 0)])
        seen = backbone.observe_rows_end()
        assert seen == {(1, 180, 180): 10283}
        assert backbone._observing is None
        assert backbone._rows_hint(t) == 10283
        assert backbone._rows_hint(_FakeTable(1, 180, 180, 9000, 1)) == 9000      # capacity clamps the hint
        assert backbone._rows_hint(_FakeTable(2, 180, 180, 64800, 1)) == 2 * 180 * 180 // 4   # other raster: guess
    finally:
        backbone._observed_rows.clear()
        backbone._observed_rows.update(saved)


def test_transposed_conv_gemm_weight_layout_matches_torch():
    """layers.dense_deconv2x2's GEMM form: weight rows (dy*2+dx)*Cout + o over K = Cin (layers.lower_deconv_gemm's
    rearrangement of layers.weight_matrix) reproduces ConvTranspose2d(k=2, s=2) (necks/rpn.py:150-154)."""
    from pillarnet_lts_b200.layers import weight_matrix
    torch.manual_seed(0)
    cin, cout, B, H, W = 8, 6, 2, 5, 7
    conv = torch.nn.ConvTranspose2d(cin, cout, 2, stride=2, bias=False)
    x = torch.randn(B, cin, H, W)
    with torch.no_grad():
        want = conv(x)
        w = weight_matrix(conv).float().reshape(cout, 4, cin).permute(1, 0, 2).reshape(4 * cout, cin)
        y = x.permute(0, 2, 3, 1).reshape(-1, cin) @ w.t()                       # (B*H*W, 4*Cout): one GEMM
        y = y.view(B, H, W, 2, 2, cout)                                          # (dy, dx) = tap >> 1, tap & 1
        got = y.permute(0, 5, 1, 3, 2, 4).reshape(B, cout, 2 * H, 2 * W)         # out(2y+dy, 2x+dx)
    assert torch.allclose(got, want, atol=1e-5)
    assert torch.allclose(F.conv_transpose2d(x, conv.weight, stride=2), want)


def test_every_reference_pillarnet_config_builds_unchanged():
    """Drop-in contract: each of the reference's configs/pillarnet/*.py is loaded with the repo's
    Config.fromfile and built through the registries without edits (class names, constructor kwargs, attr-dict task
    entries).  Only runs where the reference tree exists (the build container); parameter counts are the reference's
    (PillarNet-18 nuScenes 14,766,342 in 570 tensors)."""
    import glob
    import os
    import pytest
    cfg_dir = "/root/reference/configs/pillarnet"
    if not os.path.isdir(cfg_dir):
        pytest.skip("reference tree not present")
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200.registry import Config
    files = sorted(glob.glob(os.path.join(cfg_dir, "*.py")))
    assert len(files) == 7
    want_types = {
        "pillarnet_centerhead_nusc.py": ("PillarResNet18", "RPNV1"),
        "pillarnet_centerhead_waymo.py": ("PillarResNet18", "RPNV1"),
        "pillarnet_fpn_centerhead_waymo.py": ("PillarResNet18", "RPNG"),
        "pillarnet_fpn_iou_centerhead_waymo.py": ("PillarResNet18", "RPNG"),
        "pillarnet34_fpn_centerhead_waymo.py": ("PillarResNet34", "RPNG"),
        "pillarnet_centerhead_s4_waymo.py": ("PillarResNet18S", "RPNV2"),
        "pillarnet34_centerhead_s4_waymo.py": ("PillarResNet34S", "RPNV2"),
    }
    for f in files:
        cfg = Config.fromfile(f)
        model = P.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
        bb, nk = want_types[os.path.basename(f)]
        assert type(model.backbone).__name__ == bb and type(model.neck).__name__ == nk
        n_params = sum(p.numel() for p in model.parameters())
        if os.path.basename(f) == "pillarnet_centerhead_nusc.py":
            assert n_params == 14_766_342 and len(model.state_dict()) == 570
        H, W = model.reader.height, model.reader.width
        ps, pcr = cfg.model.reader.pillar_size, cfg.model.reader.pc_range
        assert W == round((pcr[3] - pcr[0]) / ps) and H == round((pcr[4] - pcr[1]) / ps)
        # per-task regrouping of the NMS parameters happened (detectors/pillarnet.py:25)
        nms = model.test_cfg.nms
        if nms.get("use_multi_class_nms", False):
            assert [len(v) for v in nms.nms_pre_max_size] == model.bbox_head.num_classes

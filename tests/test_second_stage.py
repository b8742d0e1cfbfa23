"""Pillar R-CNN second stage (SURVEY §8 f rank 3).

CPU: the oracle restatement (oracle/second_stage_oracle.py) against the golden vectors produced by executing the
reference's own BEVStrideFeature / PointHead / RoIMIXHead / PillarRCNN.post_process (tests/golden/second_stage.npz).
GPU: the library's modules, loaded with the golden's state dicts, against the golden (fp32 and split-bf16 tensor-core
modes 1e-4 / 1e-3 rel-to-max, bf16 mode 3e-2), the k = s sparse lateral conv and its rulebook against the oracle, and
the whole PillarRCNN detector on a toy grid."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "second_stage.npz")
CHANS = {"conv1": 32, "conv2": 64, "conv3": 128, "conv4": 256}
STRIDES = {"conv1": 1, "conv2": 2, "conv3": 4, "conv4": 8}
VARIANTS = {"a": dict(feature_sources=["conv3"], out_stride=4), "b": dict(feature_sources=["conv2", "conv3"], out_stride=2)}


def _gold():
    d = np.load(GOLD)
    return {k: torch.from_numpy(d[k]) if d[k].ndim else d[k] for k in d.files}


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(1.0, b.float().abs().max().item())


def _sd(g, prefix):
    return {k[len(prefix):]: (v if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in g.items() if k.startswith(prefix)}


def _bn(sd, name):
    return sd[name + ".weight"], sd[name + ".bias"], sd[name + ".running_mean"], sd[name + ".running_var"]


def _fc_layers(sd, idx_pairs, last):
    """[(conv index, bn index)...], last conv index -> layer dicts for the oracle"""
    layers = [dict(weight=sd[f"{c}.weight"].reshape(sd[f"{c}.weight"].shape[0], -1), bn=_bn(sd, str(b)), relu=True)
              for c, b in idx_pairs]
    if last is not None:
        layers.append(dict(weight=sd[f"{last}.weight"].reshape(sd[f"{last}.weight"].shape[0], -1), bias=sd[f"{last}.bias"]))
    return layers


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_the_reference_golden(tag):
    from oracle import second_stage_oracle as O
    g = _gold()
    ss = _sd(g, f"{tag}.second_stage.")
    v = VARIANTS[tag]
    feats = {"conv2": g["conv2"], "conv3": g["conv3"]}
    parts = [O.deconv_ks(g["bev"], ss["top_down_conv.0.weight"], _bn(ss, "top_down_conv.1"))]
    for k, src in enumerate(v["feature_sources"]):
        parts.append(O.deconv_ks(feats[src], ss[f"lat_conv.{k}.0.weight"], _bn(ss, f"lat_conv.{k}.1")))
    x = torch.cat(parts, 1)
    fused = torch.relu(O._bn(torch.nn.functional.conv2d(x, ss["fusion_conv.0.weight"], ss["fusion_conv.0.bias"], padding=1),
                             *_bn(ss, "fusion_conv.1"), 1e-3, 1))
    pc = g["pc_range"]
    cell = float(np.float32(v["out_stride"] * float(g["pillar_size"])))
    roi_f, pts = O.roi_pool(fused, g["rois"], 7, float(pc[0]), float(pc[1]), cell)
    B, N = g["rois"].shape[:2]
    assert _rel(pts, g[f"{tag}.point_coords"]) <= 1e-6
    assert _rel(roi_f.reshape(B, N, -1), g[f"{tag}.roi_features"]) <= 1e-5
    ph = _sd(g, f"{tag}.point_head.cls_layers.")
    pcs = torch.sigmoid(O.fc_stack(roi_f.reshape(-1, roi_f.shape[-1]), _fc_layers(ph, [(0, 1), (3, 4)], 6)))
    assert _rel(pcs, g[f"{tag}.point_cls_scores"]) <= 1e-5
    rh = _sd(g, f"{tag}.roi_head.")
    shared = O.fc_stack(roi_f.reshape(B * N, -1), _fc_layers(_sd(rh, "shared_fc_layer."), [(0, 1), (4, 5)], None))
    cls = O.fc_stack(shared, _fc_layers(_sd(rh, "cls_layers."), [(0, 1), (4, 5)], 7))
    reg = O.fc_stack(shared, _fc_layers(_sd(rh, "reg_layers."), [(0, 1), (4, 5)], 7))
    assert _rel(cls.view(B, N, 1), g[f"{tag}.batch_cls_preds"]) <= 1e-5
    boxes, scores, valid = O.refine(g["rois"], reg, cls, g["roi_scores"], g["roi_labels"])
    assert _rel(boxes, g[f"{tag}.batch_box_preds"]) <= 1e-5
    for i in range(B):
        assert int(valid[i].sum()) == g[f"{tag}.det{i}.scores"].shape[0]
        assert _rel(boxes[i][valid[i]], g[f"{tag}.det{i}.box3d_lidar"]) <= 1e-5
        assert _rel(scores[i][valid[i]], g[f"{tag}.det{i}.scores"]) <= 1e-5
        assert torch.equal(g["roi_labels"][i][valid[i]] - 1, g[f"{tag}.det{i}.label_preds"])


def test_pillar_rcnn_config_builds_with_the_reference_module_tree():
    """configs/pillarrcnn/pillarrcnn_fpn_centerhead_waymo.py, restated (the GPU box has no /root/reference)"""
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import configs
    cfg = configs.get("pillarrcnn_waymo")
    m = P.build_detector(cfg["model"], train_cfg=cfg["train_cfg"], test_cfg=cfg["test_cfg"])
    sd = m.state_dict()
    assert type(m).__name__ == "PillarRCNN" and len(sd) == 562
    assert sum(p.numel() for p in m.parameters()) == 16412188
    for k in ("second_stage.0.top_down_conv.0.weight", "second_stage.0.lat_conv.0.0.weight",
              "second_stage.0.fusion_conv.0.weight", "point_head.cls_layers.6.bias", "roi_head.shared_fc_layer.4.weight",
              "roi_head.reg_layers.7.weight", "single_det.bbox_head.task_heads.1.hm.3.bias"):
        assert k in sd, k
    if os.path.isdir("/root/reference/configs/pillarrcnn"):
        from pillarnet_lts_b200.registry import Config
        ref = Config.fromfile("/root/reference/configs/pillarrcnn/pillarrcnn_fpn_centerhead_waymo.py")
        m2 = P.build_detector(ref.model, train_cfg=ref.train_cfg, test_cfg=ref.test_cfg)
        assert list(m2.state_dict()) == list(sd)


# ---------------------------------------------------------------------------------------------------------------------
def _modules(tag, g, dev):
    from pillarnet_lts_b200.registry import build_point_head, build_roi_head, build_second_stage_module
    mcfg = dict(CLASS_AGNOSTIC=True, SHARED_FC=[32, 32], CLS_FC=[32, 32], REG_FC=[32, 32], DP_RATIO=0.3)
    pcfg = dict(CLASS_AGNOSTIC=True, CLS_FC=[32, 32])
    mod = build_second_stage_module(dict(type="BEVStrideFeature", grid_size=7, in_channels=128, share_channels=64,
                                         pillar_size=float(g["pillar_size"]), pc_range=[float(v) for v in g["pc_range"]],
                                         backbone_channels=CHANS, backbone_strides=STRIDES, **VARIANTS[tag]))
    head = build_roi_head(dict(type="RoIMIXHead", in_channels=64, model_cfg=mcfg, num_class=1, code_size=7, mixer_type="",
                               num_patches=49))
    phead = build_point_head(dict(type="PointHead", in_channels=64, num_class=1, model_cfg=pcfg))
    mod.load_state_dict(_sd(g, f"{tag}.second_stage."))
    head.load_state_dict(_sd(g, f"{tag}.roi_head."))
    phead.load_state_dict(_sd(g, f"{tag}.point_head."))
    return mod.to(dev).eval(), phead.to(dev).eval(), head.to(dev).eval()


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16x3", 1e-3), ("bf16", 3e-2)])
@pytest.mark.parametrize("tag", ["a", "b"])
def test_second_stage_modules_match_the_reference_golden(tag, precision, tol):
    import pillarnet_lts_b200 as P
    g = _gold()
    dev = torch.device("cuda")
    P.set_precision(precision)
    try:
        mod, phead, head = _modules(tag, g, dev)
        ex = {"rois": g["rois"].to(dev), "roi_scores": g["roi_scores"].to(dev), "roi_labels": g["roi_labels"].to(dev),
              "bev_feature": g["bev"].to(dev), "backbone_features": {"conv2": g["conv2"].to(dev), "conv3": g["conv3"].to(dev)},
              "batch_size": 2, "metadata": [None, None]}
        with torch.no_grad():
            ex = mod(ex)
            ex = phead(ex)
            out = head(ex, training=False)
        torch.cuda.synchronize()
    finally:
        P.set_precision("bf16")
    assert _rel(ex["point_coords"].cpu(), g[f"{tag}.point_coords"]) <= 1e-5
    assert _rel(ex["roi_features"].cpu(), g[f"{tag}.roi_features"]) <= tol
    assert _rel(ex["point_cls_scores"].cpu(), g[f"{tag}.point_cls_scores"]) <= tol
    assert _rel(out["batch_cls_preds"].cpu(), g[f"{tag}.batch_cls_preds"]) <= tol
    assert _rel(out["batch_box_preds"].cpu(), g[f"{tag}.batch_box_preds"]) <= tol
    for i in range(2):
        m = out["refined_valid"][i].cpu()
        assert int(m.sum()) == g[f"{tag}.det{i}.scores"].shape[0]
        assert _rel(out["refined_scores"][i].cpu()[m], g[f"{tag}.det{i}.scores"]) <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("s", [1, 2, 4])
def test_block_rulebook_and_sparse_lateral_conv_match_the_oracle(s):
    """pn_rulebook_block (k = s, stride = s) bit-exact vs a numpy restatement; SparseConv2d(k = s) + BN1d + ReLU through
    it vs the dense-equivalent oracle (fp32 1e-4; bf16 3e-2)"""
    import pillarnet_lts_b200 as P
    from oracle import second_stage_oracle as O
    from pillarnet_lts_b200 import ops
    from pillarnet_lts_b200.layers import SparseConv2d, SparseSequential, build_norm_layer
    from pillarnet_lts_b200.second_stage import sparse_block_conv
    from pillarnet_lts_b200.sparse import SparseConvTensor
    rng = np.random.default_rng(40 + s)
    dev = torch.device("cuda")
    B, H, W, cin, cout = 2, 37, 50, 32, 64
    act = rng.random((B, H, W)) < 0.15
    act[0, :4, :4] = True
    coords = np.argwhere(act).astype(np.int32)
    n = coords.shape[0]
    pts = np.zeros((n, 5), np.float32)
    pts[:, 0] = coords[:, 2] + 0.5
    pts[:, 1] = coords[:, 1] + 0.5
    offs = np.cumsum([0] + [int((coords[:, 0] == b).sum()) for b in range(B)]).astype(np.int32)
    table, _ = ops.pillarize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), B, H, W, 0.0, 0.0, 1.0)
    assert table.count() == n and np.array_equal(table.coords[:n].cpu().numpy(), coords)
    out_table, nbr = ops.rulebook_block(table, s)
    Ho, Wo = H // s, W // s
    blk = act[:, :Ho * s, :Wo * s].reshape(B, Ho, s, Wo, s).any(axis=(2, 4))
    want_coords = np.argwhere(blk).astype(np.int32)
    m = out_table.count()
    assert m == want_coords.shape[0] and np.array_equal(out_table.coords[:m].cpu().numpy(), want_coords)
    rank = -np.ones((B, H, W), np.int64)
    rank[act] = np.arange(n)
    want_nbr = np.stack([rank[want_coords[:, 0], want_coords[:, 1] * s + k // s, want_coords[:, 2] * s + k % s]
                         for k in range(s * s)], 1)
    assert np.array_equal(nbr[:m].cpu().numpy(), want_nbr)
    for precision, tol in (("fp32", 1e-4), ("bf16", 3e-2)):
        P.set_precision(precision)
        try:
            torch.manual_seed(s)
            seq = SparseSequential(SparseConv2d(cin, cout, kernel_size=s, stride=s, padding=0, bias=True),
                                   build_norm_layer(dict(type="BN1d", eps=1e-3, momentum=0.01), cout)[1],
                                   torch.nn.ReLU()).to(dev).eval()
            seq[1].running_mean.normal_(0, 0.1)
            seq[1].running_var.uniform_(0.5, 1.5)
            feat = torch.zeros(table.cap, cin, device=dev)
            feat[:n] = torch.randn(n, cin, device=dev)
            if precision == "bf16":
                feat = feat.to(torch.bfloat16)
            sp = SparseConvTensor(feat, table, (H, W), B)
            with torch.no_grad():
                got = sparse_block_conv(sp, seq).dense().float().cpu()
            x = torch.zeros(B, cin, H, W)
            x[coords[:, 0], :, coords[:, 1], coords[:, 2]] = feat[:n].float().cpu()
            bn = seq[1]
            want = O.block_sparse_conv(x, torch.from_numpy(act), seq[0].weight.detach().cpu(), seq[0].bias.detach().cpu(),
                                       (bn.weight.detach().cpu(), bn.bias.detach().cpu(), bn.running_mean.cpu(),
                                        bn.running_var.cpu()))
            assert _rel(got, want) <= tol, precision
        finally:
            P.set_precision("bf16")


@pytest.mark.gpu
def test_roi_grid_bilinear_on_a_padded_bf16_map_and_far_away_rois():
    """the padded (zero-bordered) bf16 layout and RoIs far outside the map (clamped corners, weights that do not sum
    to one — as the reference computes them) against the oracle"""
    from oracle import second_stage_oracle as O
    from pillarnet_lts_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, H, W, C, N = 2, 19, 23, 64, 9
    fmap = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16)
    padded = torch.zeros(B, H + 2, W + 2, C + 8, dtype=torch.bfloat16)
    padded[:, 1:-1, 1:-1, 8:] = fmap
    rois = torch.zeros(B, N, 7)
    rois[..., :2] = torch.rand(B, N, 2, generator=g) * 30 - 4
    rois[..., 3:5] = torch.rand(B, N, 2, generator=g) * 4 + 0.2
    rois[..., 6] = torch.rand(B, N, generator=g) * 7 - 3.5
    rois[0, 0, :2] = torch.tensor([1e4, -1e4])
    want, want_pts = O.roi_pool(fmap.float().permute(0, 3, 1, 2), rois, 5, -2.0, -1.0, 1.25)
    got, pts = ops.roi_grid_bilinear(rois.cuda(), 5, padded.view(-1, C + 8).cuda(), B, H, W, C, -2.0, -1.0, 1.25,
                                     feat_coff=8, padded=True)
    torch.cuda.synchronize()
    assert _rel(pts.cpu(), want_pts) <= 1e-6
    assert _rel(got.float().cpu(), want) <= 1e-2        # one bf16 ulp of the output


@pytest.mark.gpu
def test_pillar_rcnn_detector_runs_end_to_end_and_refines_the_first_stage_boxes():
    """toy-grid PillarRCNN (PillarResNet18 + RPNG + CenterHead at strides 8 / 4 + BEVStrideFeature + PointHead +
    RoIMIXHead): the detector's output equals the second-stage oracle applied to the detector's own first-stage
    outputs (fp32 mode, 1e-4), and the bf16 mode runs"""
    import pillarnet_lts_b200 as P
    from oracle import second_stage_oracle as O
    from pillarnet_lts_b200 import configs, synth
    cfg = configs.get("pillarrcnn_toy")
    dev = torch.device("cuda")
    torch.manual_seed(3)
    model = P.build_detector(cfg["model"], train_cfg=None, test_cfg=cfg["test_cfg"]).to(dev).eval()
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.normal_(0, 0.05)
            m.running_var.uniform_(0.8, 1.2)
    frames = [torch.from_numpy(f).to(dev) for f in synth.make_batch(cfg["synth"], 2, 7)]
    for precision in ("fp32", "bf16"):
        P.set_precision(precision)
        try:
            with torch.no_grad():
                ex = dict(points=frames, metadata=[{}, {}])
                dets = model(ex, return_loss=False)
            torch.cuda.synchronize()
            assert len(dets) == 2
            for d in dets:
                assert d["box3d_lidar"].shape[1] == 7 and d["scores"].shape[0] == d["box3d_lidar"].shape[0]
            # the sync-free path (every NMS slot a RoI, validity on the device) returns the same detections
            offs = torch.tensor([0, frames[0].shape[0], frames[0].shape[0] + frames[1].shape[0]], dtype=torch.int32, device=dev)
            boxes_d, scores_d, labels_d, valid_d = model.forward_device(torch.cat(frames), offs)
            torch.cuda.synchronize()
            for i, d in enumerate(dets):
                m = valid_d[i]
                assert int(m.sum()) == d["scores"].shape[0]
                got = torch.cat([boxes_d[i][m], scores_d[i][m][:, None], labels_d[i][m][:, None].float()], 1)
                want = torch.cat([d["box3d_lidar"], d["scores"][:, None], d["label_preds"][:, None].float()], 1)
                # same rows in a different order: sort both lexicographically by (label, score, x)
                def canon(t):
                    key = t[:, -1] * 1e6 + t[:, -2] * 1e3 + t[:, 0] * 1e-3
                    return t[torch.argsort(key)]
                assert _rel(canon(got), canon(want)) <= (1e-5 if precision == "fp32" else 1e-3), (precision, i)
            if precision != "fp32":
                continue
            # oracle on the same first-stage outputs
            ss = model.second_stage[0]
            sd = {k: v.detach().cpu() for k, v in ss.state_dict().items()}
            bev = ex["bev_feature"].float().cpu()
            c3 = ex["backbone_features"]["conv3"].dense().float().cpu()
            parts = [O.deconv_ks(bev, sd["top_down_conv.0.weight"], _bn(sd, "top_down_conv.1")),
                     O.deconv_ks(c3, sd["lat_conv.0.0.weight"], _bn(sd, "lat_conv.0.1"))]
            fused = torch.relu(O._bn(torch.nn.functional.conv2d(torch.cat(parts, 1), sd["fusion_conv.0.weight"],
                                                                sd["fusion_conv.0.bias"], padding=1),
                                     *_bn(sd, "fusion_conv.1"), 1e-3, 1))
            rois = ex["rois"].cpu()
            pcr = ss.point_cloud_range
            roi_f, _ = O.roi_pool(fused, rois, ss.grid_size, pcr[0], pcr[1], float(np.float32(ss.out_stride * ss.pillar_size)))
            B, N = rois.shape[:2]
            assert ex["rois"].abs().sum() > 0 and _rel(ex["roi_features"].cpu(), roi_f.reshape(B, N, -1)) <= 1e-4
            rh = {k: v.detach().cpu() for k, v in model.roi_head.state_dict().items()}
            shared = O.fc_stack(roi_f.reshape(B * N, -1), _fc_layers(_sd(rh, "shared_fc_layer."), [(0, 1), (4, 5)], None))
            cls = O.fc_stack(shared, _fc_layers(_sd(rh, "cls_layers."), [(0, 1), (4, 5)], 7))
            reg = O.fc_stack(shared, _fc_layers(_sd(rh, "reg_layers."), [(0, 1), (4, 5)], 7))
            boxes, scores, valid = O.refine(rois, reg, cls, ex["roi_scores"].cpu(), ex["roi_labels"].cpu())
            for i in range(B):
                assert int(valid[i].sum()) == dets[i]["scores"].shape[0] > 0
                assert _rel(dets[i]["box3d_lidar"].cpu(), boxes[i][valid[i]]) <= 1e-4
                assert _rel(dets[i]["scores"].cpu(), scores[i][valid[i]]) <= 1e-4
        finally:
            P.set_precision("bf16")

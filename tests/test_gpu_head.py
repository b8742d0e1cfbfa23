"""GPU parity of the dense neck/head convs and CenterHead.predict.

neck/head forward: golden vectors from the reference's RPNV1 + CenterHead executed with torch (CPU,
fp32) in the build container; fp32 mode tolerance max-abs 1e-3 relative to max|ref| (north_star).
predict: golden vectors from the reference's CenterHead.predict (circular NMS); rotated path against a
torch-CUDA restatement of center_head.py:257-413 + the reference's nms_gpu from oracle/_ref."""
import logging
import os

import numpy as np
import pytest
import torch

from oracle import pillarnet_oracle as O
from tests.gpu_util import ref_ext

pytestmark = pytest.mark.gpu

TASKS = [dict(stride=8, class_names=["car"]), dict(stride=8, class_names=["ped", "cone"])]
HEADS = {"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)}
PS, PCR = 0.075, [-54, -54, -5.0, 54, 54, 3.0]


def _load_state(module, g, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}
    missing, unexpected = module.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def test_neck_head_forward_vs_reference_golden(golden_dir):
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200.head import CenterHead
    from pillarnet_lts_b200.neck import RPNV1
    g = np.load(os.path.join(golden_dir, "neck_head_forward.npz"))
    P.set_precision("fp32")
    neck = RPNV1(layer_nums=[1, 2], num_filters=32, in_channels=[32, 32], logger=logging.getLogger("t"))
    head = CenterHead(tasks=TASKS, in_channels=[32], code_weights=[1.0] * 10, common_heads=HEADS,
                      share_channel=16, pillar_size=PS, point_cloud_range=PCR)
    _load_state(neck, g, "neck.")   # state_dict keys/layouts are the reference's (strict load)
    _load_state(head, g, "head.")
    neck.cuda().eval()
    head.cuda().eval()
    x4, x5 = torch.from_numpy(g["x4"]).cuda(), torch.from_numpy(g["x5"]).cuda()
    bev = neck({"conv4": x4, "conv5": x5})
    preds = head(bev)
    torch.cuda.synchronize()
    want = g["bev"]
    assert np.abs(bev[0].float().cpu().numpy() - want).max() <= 1e-3 * max(1.0, np.abs(want).max())
    for t, p in enumerate(preds):
        for k, v in p.items():
            w = g[f"pred{t}_{k}"]
            assert v.shape == w.shape
            assert np.abs(v.cpu().numpy() - w).max() <= 1e-3 * max(1.0, np.abs(w).max()), (t, k)


def _golden_preds(g):
    preds = []
    for t in range(2):
        preds.append({n: torch.from_numpy(g[f"t{t}_{n}"]).cuda() for n in ["reg", "height", "dim", "rot", "vel", "hm"]})
    return preds


def test_predict_circle_vs_reference_golden(golden_dir):
    from pillarnet_lts_b200.head import CenterHead
    from pillarnet_lts_b200.registry import ConfigDict
    g = np.load(os.path.join(golden_dir, "head_predict_circle.npz"))
    head = CenterHead(tasks=TASKS, in_channels=[16], code_weights=[1.0] * 10, common_heads=HEADS,
                      share_channel=8, pillar_size=PS, point_cloud_range=PCR).cuda()
    cfg = ConfigDict.wrap(dict(circular_nms=True, min_radius=[4.0, 0.85],
                               nms=dict(nms_pre_max_size=[1000, 1000], nms_post_max_size=[83, 83],
                                        nms_iou_threshold=0.2),
                               score_threshold=0.1, post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]))
    rets = head.predict({"metadata": [None, None]}, _golden_preds(g), cfg)
    for b, r in enumerate(rets):
        assert np.array_equal(r["label_preds"].cpu().numpy(), g[f"out{b}_labels"])
        np.testing.assert_allclose(r["scores"].cpu().numpy(), g[f"out{b}_scores"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(r["box3d_lidar"].cpu().numpy(), g[f"out{b}_boxes"], rtol=1e-5, atol=1e-5)


HEADS_IOU = dict(HEADS, iou=(1, 2))


def test_predict_double_flip_vs_reference_golden(golden_dir):
    """double-flip TTA (center_head.py:233-304): 8 input frames -> 2 output frames, vs the reference's predict."""
    from pillarnet_lts_b200.head import CenterHead
    from pillarnet_lts_b200.registry import ConfigDict
    g = np.load(os.path.join(golden_dir, "head_predict_double_flip.npz"))
    head = CenterHead(tasks=TASKS, in_channels=[16], code_weights=[1.0] * 10, common_heads=HEADS_IOU,
                      share_channel=8, pillar_size=PS, point_cloud_range=PCR).cuda()
    cfg = ConfigDict.wrap(dict(circular_nms=True, min_radius=[4.0, 0.85], double_flip=True,
                               nms=dict(nms_pre_max_size=[1000, 1000], nms_post_max_size=[83, 83],
                                        nms_iou_threshold=0.2),
                               score_threshold=0.1, post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]))
    preds = [{n: torch.from_numpy(g[f"t{t}_{n}"]).cuda() for n in ["reg", "height", "dim", "rot", "vel", "iou", "hm"]}
             for t in range(2)]
    rets = head.predict({"metadata": list(range(8))}, preds, cfg)
    assert len(rets) == 2 and [r["metadata"] for r in rets] == [0, 4]
    for b, r in enumerate(rets):
        assert np.array_equal(r["label_preds"].cpu().numpy(), g[f"out{b}_labels"])
        np.testing.assert_allclose(r["scores"].cpu().numpy(), g[f"out{b}_scores"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(r["box3d_lidar"].cpu().numpy(), g[f"out{b}_boxes"], rtol=1e-5, atol=1e-5)


def test_double_flip_merge_bit_exact_vs_torch():
    """pn_double_flip_merge vs the torch-CUDA op sequence of center_head.py:233-304 on the same maps: bit-exact."""
    from pillarnet_lts_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    Bo, H, W, K = 3, 33, 47, 2
    offs = {"reg": 0, "height": 2, "dim": 3, "rot": 6, "vel": 8, "iou": 10, "hm": 11}
    C = 11 + K
    maps = torch.randn(Bo * 4, H, W, C, device="cuda", generator=g) * 1.5
    wide = torch.full((Bo * 4 * H * W, C + 5), 7.0, device="cuda")
    wide[:, 2:2 + C] = maps.view(-1, C)
    got = ops.double_flip_merge(wide[:, 2:2 + C], offs, K, Bo, H, W).view(Bo, H, W, C)
    v = maps.clone().view(Bo, 4, H, W, C)
    v[:, 1] = torch.flip(v[:, 1], dims=[1])
    v[:, 2] = torch.flip(v[:, 2], dims=[2])
    v[:, 3] = torch.flip(v[:, 3], dims=[1, 2])
    hm = torch.sigmoid(v[..., 11:]).mean(dim=1)
    dim = torch.exp(v[..., 3:6].clamp(min=-1.2, max=3.2)).mean(dim=1)
    iou = torch.clamp((v[..., 10] + 1) * 0.5, min=0, max=1.).mean(dim=1)
    reg, rots, rotc, vel = v[..., 0:2], v[..., 6:7], v[..., 7:8], v[..., 8:10]
    reg[:, 1, ..., 1] = 1 - reg[:, 1, ..., 1]
    reg[:, 2, ..., 0] = 1 - reg[:, 2, ..., 0]
    reg[:, 3, ..., 0] = 1 - reg[:, 3, ..., 0]
    reg[:, 3, ..., 1] = 1 - reg[:, 3, ..., 1]
    rotc[:, 1] *= -1
    rots[:, 2] *= -1
    rots[:, 3] *= -1
    rotc[:, 3] *= -1
    vel[:, 1, ..., 1] *= -1
    vel[:, 2, ..., 0] *= -1
    vel[:, 3] *= -1
    want = torch.cat([reg.mean(dim=1), v[..., 2:3].mean(dim=1), dim, rots.mean(dim=1), rotc.mean(dim=1),
                      vel.mean(dim=1), iou.unsqueeze(-1), hm], dim=-1)
    torch.cuda.synchronize()
    for name, lo, hi in (("reg", 0, 2), ("height", 2, 3), ("dim", 3, 6), ("rot", 6, 8), ("vel", 8, 10),
                         ("iou", 10, 11), ("hm", 11, C)):
        assert torch.equal(got[..., lo:hi], want[..., lo:hi]), name


def _torch_predict_rotate(preds, strides, num_classes, cfg, nms_gpu):
    """center_head.py:216-413 + box_torch_ops.py:296-322 restated with torch CUDA ops (stable sort)."""
    outs = None
    rets = []
    for t, p in enumerate(preds):
        p = {k: v.permute(0, 2, 3, 1).contiguous() for k, v in p.items()}
        hm = torch.sigmoid(p["hm"])
        dim = torch.exp(p["dim"].clamp(min=-1.2, max=3.2))
        rot = torch.atan2(p["rot"][..., 0:1], p["rot"][..., 1:2])
        B, H, W, _ = hm.shape
        ys, xs = torch.meshgrid([torch.arange(0, H), torch.arange(0, W)], indexing="ij")
        ys = ys.view(1, H, W).repeat(B, 1, 1).to(hm)
        xs = xs.view(1, H, W).repeat(B, 1, 1).to(hm)
        xs = xs.view(B, H, W, 1) + p["reg"][..., 0:1]
        ys = ys.view(B, H, W, 1) + p["reg"][..., 1:2]
        xs = xs * strides[t] * PS + PCR[0]
        ys = ys * strides[t] * PS + PCR[1]
        boxes = torch.cat([xs, ys, p["height"], dim, p["vel"], rot], dim=-1)
        rng = torch.tensor(cfg["post_center_limit_range"], dtype=hm.dtype, device=hm.device)
        frames = []
        for b in range(B):
            bp = boxes[b].reshape(-1, 9)
            scores, labels = torch.max(hm[b].reshape(H * W, -1), dim=-1)
            m = (scores > cfg["score_threshold"]) & (bp[:, :3] >= rng[:3]).all(-1) & (bp[:, :3] <= rng[3:]).all(-1)
            bp, scores, labels = bp[m], scores[m], labels[m]
            order = torch.sort(scores, descending=True, stable=True)[1][:cfg["pre"]]
            pc = bp[order][:, [0, 1, 2, 4, 3, 5, -1]]
            pc[:, -1] = -pc[:, -1] - np.pi / 2
            pc = pc.contiguous()
            keep = torch.LongTensor(pc.size(0))
            n = nms_gpu(pc, keep, cfg["thr"]) if pc.size(0) else 0
            sel = order[keep[:n].cuda()][:cfg["post"]]
            frames.append((bp[sel], scores[sel], labels[sel] + sum(num_classes[:t])))
        rets.append(frames)
    B = len(rets[0])
    outs = []
    for b in range(B):
        outs.append(tuple(torch.cat([r[b][i] for r in rets]) for i in range(3)))
    return outs


@pytest.mark.parametrize("n_peaks", [2500, 9000])
def test_predict_rotate_bit_exact_vs_torch_and_reference_nms(n_peaks):
    """n_peaks 2500: every NMS segment has <= 4096 candidates (8-CTA cluster selection, k_select_topk_cluster);
    9000: more than 4096 (single-CTA radix select, k_select_topk)."""
    iou3d = ref_ext("iou3d_nms_cuda")
    if iou3d is None:
        pytest.skip("oracle/_ref/iou3d_nms_cuda not built")
    from pillarnet_lts_b200 import synth
    from pillarnet_lts_b200.head import CenterHead
    from pillarnet_lts_b200.registry import ConfigDict
    rng = np.random.default_rng(31)
    B, H, W = 2, 180, 180
    head = CenterHead(tasks=TASKS, in_channels=[16], code_weights=[1.0] * 10, common_heads=HEADS,
                      share_channel=8, pillar_size=PS, point_cloud_range=PCR).cuda()
    preds = []
    for t, K in enumerate([1, 2]):
        m = synth.synthetic_head_maps(rng, B, H, W, 10 + K, slice(10, 10 + K), n_peaks=n_peaks)
        m[..., 0:2] = rng.uniform(0, 1, m[..., 0:2].shape)
        m[..., 3:6] = rng.normal(0.5, 0.5, m[..., 3:6].shape)
        tm = torch.from_numpy(m).cuda().permute(0, 3, 1, 2).contiguous()
        preds.append({"reg": tm[:, 0:2], "height": tm[:, 2:3], "dim": tm[:, 3:6], "rot": tm[:, 6:8],
                      "vel": tm[:, 8:10], "hm": tm[:, 10:]})
    cfg = ConfigDict.wrap(dict(nms=dict(use_rotate_nms=True, nms_pre_max_size=1000, nms_post_max_size=83,
                                        nms_iou_threshold=0.2), rectifier=0, score_threshold=0.1,
                               post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]))
    got = head.predict({"metadata": [None] * B}, preds, cfg)
    _, _, plan = head.predict_raw(preds, cfg)
    most = int(plan["cand_count"].max())
    assert (most > 4096) == (n_peaks > 4096), most
    want = _torch_predict_rotate(preds, [8, 8], [1, 2], dict(post_center_limit_range=cfg.post_center_limit_range,
                                                            score_threshold=0.1, pre=1000, post=83, thr=0.2),
                                 iou3d.nms_gpu)
    torch.cuda.synchronize()
    for b in range(B):
        wb, ws, wl = want[b]
        assert wb.shape[0] > 50
        assert torch.equal(got[b]["label_preds"], wl)
        assert torch.equal(got[b]["scores"], ws)        # bit-exact scores
        assert torch.equal(got[b]["box3d_lidar"], wb)   # bit-exact decoded boxes and keep list


def test_predict_multi_class_vs_oracle():
    """Waymo-style per-class NMS (rotate_class_nms_pcdet) vs the CPU oracle."""
    from pillarnet_lts_b200 import synth
    from pillarnet_lts_b200.head import CenterHead
    from pillarnet_lts_b200.registry import ConfigDict
    rng = np.random.default_rng(32)
    tasks = [dict(stride=8, class_names=["VEHICLE"]), dict(stride=4, class_names=["PEDESTRIAN", "CYCLIST"])]
    heads = {"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "iou": (1, 2)}
    ps, pcr = 0.1, [-75.2, -75.2, -2, 75.2, 75.2, 4]
    head = CenterHead(tasks=tasks, in_channels=[16, 8], code_weights=[1.0] * 8, common_heads=heads,
                      share_channel=8, pillar_size=ps, point_cloud_range=pcr).cuda()
    B = 2
    preds, raw = [], []
    for t, (K, HW) in enumerate([(1, 94), (2, 188)]):
        m = synth.synthetic_head_maps(rng, B, HW, HW, 9 + K, slice(9, 9 + K), n_peaks=3000)
        m[..., 0:2] = rng.uniform(0, 1, m[..., 0:2].shape)
        raw.append(m)
        tm = torch.from_numpy(m).cuda().permute(0, 3, 1, 2).contiguous()
        preds.append({"reg": tm[:, 0:2], "height": tm[:, 2:3], "dim": tm[:, 3:6], "rot": tm[:, 6:8],
                      "iou": tm[:, 8:9], "hm": tm[:, 9:]})
    cfg = ConfigDict.wrap(dict(nms=dict(use_multi_class_nms=True, nms_pre_max_size=[[2048], [1024, 1024]],
                                        nms_post_max_size=[[200], [150, 150]],
                                        nms_iou_threshold=[[0.8], [0.55, 0.55]]),
                               rectifier=[[0.0], [0.0, 0.0]], score_threshold=0.1,
                               post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0]))
    got = head.predict({"metadata": [None] * B}, preds, cfg)
    offs = {"reg": 0, "height": 2, "dim": 3, "rot": 6, "iou": 8, "hm": 9}
    for b in range(B):
        ob, os_, ol = [], [], []
        for t, (K, stride) in enumerate([(1, 8), (2, 4)]):
            boxes, hm, iou = O.decode_task(raw[t], offs, K, stride, ps, pcr)
            c = dict(mode="multi_class", rectifiers=cfg.rectifier[t], thrs=cfg.nms.nms_iou_threshold[t],
                     pre_max=cfg.nms.nms_pre_max_size[t], post_max=cfg.nms.nms_post_max_size[t],
                     score_threshold=0.1, post_center_limit_range=cfg.post_center_limit_range)
            bx, sc, lb = O.post_process_frame(boxes[b], hm[b], iou[b], c)
            ob.append(bx); os_.append(sc); ol.append(lb + (0 if t == 0 else 1))
        wb, ws, wl = np.concatenate(ob), np.concatenate(os_), np.concatenate(ol)
        assert len(wb) > 100
        assert np.array_equal(got[b]["label_preds"].cpu().numpy(), wl)
        np.testing.assert_allclose(got[b]["scores"].cpu().numpy(), ws, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(got[b]["box3d_lidar"].cpu().numpy(), wb, rtol=1e-5, atol=1e-5)

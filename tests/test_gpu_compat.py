"""GPU: the reference's OWN Python (staged by oracle/build_ref.py into oracle/_ref/py, git-ignored) executed on the
operator-level drop-in modules `pillarnet_lts_b200.compat.{pillar_cuda, iou3d_nms_cuda}` — a second, reference-driven
parity channel: det3d's DynamicPFE / PillarQueryAndGroup / PillarMaxPooling / scatter_max
(dynamic_pillar_encoder.py:29-50, pillar_utils.py:22-57, pillar_modules.py:56-74, scatter_utils.py:7-37) and
rotate_nms_pcdet / rotate_class_nms_pcdet / boxes_iou3d_gpu (box_torch_ops.py:296-359, iou3d_nms_utils.py:37-71)
run unmodified, once bound to the reference's compiled extensions (oracle/_ref/*.so) and once bound to this library,
on identical inputs; and the fused product path (DynamicPFE here) is compared with both.

Bit-exact: pillar indices, point->pillar indices, point features, keep lists, selected boxes.  scatter-max outputs
bit-exact; its gradient equal except at exact-tie/near-tie argmax (the reference routes to any point within 1e-5).
"""
import importlib
import sys

import numpy as np
import pytest
import torch

from oracle import build_ref
from tests.gpu_util import rand_boxes, ref_ext

pytestmark = pytest.mark.gpu

PS, PCR = 0.075, [-54, -54, -5.0, 54, 54, 3.0]


@pytest.fixture(scope="module")
def det3d_ref():
    """the staged reference Python with its extension modules bound to the compat shims; returns a namespace whose
    .bind("ours"|"ref") re-points every staged module at this library or at the reference's compiled .so"""
    if not build_ref.python_available():
        pytest.skip("oracle/_ref/py not staged (run __graft_entry__.build() where /root/reference exists)")
    from pillarnet_lts_b200 import compat
    sys.path.insert(0, build_ref.PY_OUT)
    saved = {k: v for k, v in sys.modules.items() if k == "det3d" or k.startswith("det3d.") or k.startswith("spconv")}
    for k in saved:
        del sys.modules[k]
    compat.install("det3d.ops")
    mods = {n: importlib.import_module(n) for n in (
        "det3d.ops.pillar_ops.group_utils", "det3d.ops.pillar_ops.scatter_utils", "det3d.ops.pillar_ops.pillar_utils",
        "det3d.ops.pillar_ops.pillar_modules", "det3d.ops.iou3d_nms.iou3d_nms_utils", "det3d.core.bbox.box_torch_ops",
        "det3d.models.readers.dynamic_pillar_encoder")}

    class NS:
        pass

    ns = NS()
    ns.mods = mods
    ns.reader = mods["det3d.models.readers.dynamic_pillar_encoder"]
    ns.box_ops = mods["det3d.core.bbox.box_torch_ops"]
    ns.iou_utils = mods["det3d.ops.iou3d_nms.iou3d_nms_utils"]
    ns.scatter = mods["det3d.ops.pillar_ops.scatter_utils"]
    ref_pillar, ref_iou = ref_ext("pillar_cuda"), ref_ext("iou3d_nms_cuda")

    def bind(which):
        if which == "ref" and (ref_pillar is None or ref_iou is None):
            pytest.skip("reference extensions did not travel")
        pc = compat.pillar_cuda if which == "ours" else ref_pillar
        ic = compat.iou3d_nms_cuda if which == "ours" else ref_iou
        for n in ("group_utils", "scatter_utils", "pillar_utils"):
            mods[f"det3d.ops.pillar_ops.{n}"].pillar_cuda = pc
        ns.iou_utils.iou3d_nms_cuda = ic
        ns.box_ops.iou3d_nms_cuda = ic

    ns.bind = bind
    yield ns
    sys.path.remove(build_ref.PY_OUT)
    for k in [k for k in sys.modules if k == "det3d" or k.startswith("det3d.") or k.startswith("spconv")]:
        del sys.modules[k]
    sys.modules.update(saved)


def _frames(n, seed0=40):
    from pillarnet_lts_b200 import synth
    return [torch.from_numpy(synth.make_frame("nuscenes", seed0 + i)[::2].copy()).cuda() for i in range(n)]


def _run_reference_reader(ns, pts, weight_from):
    torch.manual_seed(0)
    rd = ns.reader.DynamicPFE(in_channels=5, num_filters=(32,), pillar_size=PS, pc_range=PCR).cuda().eval()
    rd.pfn_layers.load_state_dict(weight_from.pfn_layers.state_dict(), strict=False)
    with torch.no_grad():
        sp = rd(dict(points=pts))
    return rd, sp


def test_reference_reader_python_runs_on_the_compat_module_and_matches_both(det3d_ref):
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200.reader import DynamicPFE
    from tests.gpu_util import randomize_bn
    P.set_precision("fp32")
    torch.manual_seed(1)
    ours = DynamicPFE(in_channels=5, num_filters=(32,), pillar_size=PS, pc_range=PCR).cuda().eval()
    randomize_bn(ours, 3)
    pts = _frames(2)
    out = {}
    for which in ("ours", "ref"):
        det3d_ref.bind(which)
        _, sp = _run_reference_reader(det3d_ref, pts, ours)
        out[which] = (sp.indices.clone(), sp.features.clone())
        assert sp.spatial_shape == (1440, 1440) and sp.batch_size == 2
    torch.cuda.synchronize()
    # the reference's Python gives the same pillars and features whichever extension it calls
    assert torch.equal(out["ours"][0], out["ref"][0])
    assert torch.equal(out["ours"][1], out["ref"][1])
    # and the fused product path reproduces them: indices bit-exact, features to fp32 rounding of the folded BN
    with torch.no_grad():
        spf = ours(dict(points=pts))
    assert torch.equal(spf.indices, out["ref"][0])
    ref_f = out["ref"][1]
    err = (spf.features_f32[: ref_f.shape[0]] - ref_f).abs().max().item() / max(1.0, ref_f.abs().max().item())
    assert err <= 1e-4, err


def test_reference_group_python_intermediates_bit_exact(det3d_ref):
    """PillarQueryAndGroup.forward (pillar_utils.py:22-57) step by step on both extensions"""
    pu = det3d_ref.mods["det3d.ops.pillar_ops.pillar_utils"]
    pts = _frames(3, seed0=50)
    xy, cnt, feats = [], [], []
    for p in pts:
        cx = ((p[:, 0] - PCR[0]) / PS).floor().int()
        cy = ((p[:, 1] - PCR[1]) / PS).floor().int()
        m = (cx >= 0) & (cx < 1440) & (cy >= 0) & (cy < 1440)
        xy.append(torch.stack((cx[m], cy[m]), 1))
        feats.append(p[m])
        cnt.append(int(m.sum()))
    xy, feats = torch.cat(xy), torch.cat(feats)
    cnt = torch.tensor(cnt, dtype=torch.int32, device="cuda")
    res = {}
    for which in ("ours", "ref"):
        det3d_ref.bind(which)
        g = pu.PillarQueryAndGroup(PS, PCR)
        res[which] = g(xy, cnt, feats)
    for a, b in zip(res["ours"], res["ref"]):
        assert a.dtype == b.dtype and torch.equal(a, b)


def test_reference_scatter_max_autograd_on_the_compat_module(det3d_ref):
    rng = np.random.default_rng(5)
    L, M, C = 50_000, 9_000, 32
    idx = torch.from_numpy(rng.integers(0, M, L).astype(np.int32)).cuda()
    src0 = torch.from_numpy(rng.standard_normal((L, C)).astype(np.float32)).cuda()
    gout = torch.from_numpy(rng.standard_normal((M, C)).astype(np.float32)).cuda()
    res = {}
    for which in ("ours", "ref"):
        det3d_ref.bind(which)
        src = src0.clone().requires_grad_(True)
        out = det3d_ref.scatter.scatter_max(src, idx, M)
        out.backward(gout)
        res[which] = (out.detach(), src.grad.clone())
    assert torch.equal(res["ours"][0], res["ref"][0])
    # gradient: identical routing except where two points of a pillar tie within the reference's 1e-5 window
    diff = (res["ours"][1] != res["ref"][1]).any(1).float().mean().item()
    assert diff <= 1e-3, diff
    assert torch.allclose(res["ours"][1].sum(0), res["ref"][1].sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("n,clusters", [(700, 25), (2048, 60)])
def test_reference_rotate_nms_python_on_the_compat_module(det3d_ref, n, clusters):
    rng = np.random.default_rng(n)
    b7 = rand_boxes(rng, n, spread=30.0, clusters=clusters)
    # det3d box layout [x,y,z,w,l,h,vx,vy,rot]
    boxes = np.zeros((n, 9), np.float32)
    boxes[:, :6] = b7[:, :6]
    boxes[:, 8] = b7[:, 6]
    boxes = torch.from_numpy(boxes).cuda()
    scores = torch.from_numpy(rng.random(n).astype(np.float32)).cuda()
    ious = torch.from_numpy(rng.random(n).astype(np.float32)).cuda()
    labels = torch.from_numpy(rng.integers(0, 3, n)).cuda()
    res = {}
    for which in ("ours", "ref"):
        det3d_ref.bind(which)
        single = det3d_ref.box_ops.rotate_nms_pcdet(boxes, scores, ious, labels, rectifier=0.5, nms_thresh=0.2,
                                                    pre_maxsize=1000, post_max_size=83, use_rectify=True)
        multi = det3d_ref.box_ops.rotate_class_nms_pcdet(boxes, scores, ious, labels, nms_thresh=[0.8, 0.55, 0.55],
                                                         rectifiers=[0.68, 0.71, 0.65], pre_maxsize=[2048, 1024, 1024],
                                                         post_max_size=[200, 150, 150])
        iou3d = det3d_ref.iou_utils.boxes_iou3d_gpu(boxes[:200, [0, 1, 2, 3, 4, 5, 8]].contiguous(),
                                                   boxes[100:400, [0, 1, 2, 3, 4, 5, 8]].contiguous())
        res[which] = (single, multi, iou3d)
    for a, b in zip(res["ours"][0] + res["ours"][1], res["ref"][0] + res["ref"][1]):
        assert a.shape == b.shape and torch.equal(a, b)
    assert torch.equal(res["ours"][2], res["ref"][2])


def test_compat_nms_normal_matches_reference_extension(det3d_ref):
    ref = ref_ext("iou3d_nms_cuda")
    if ref is None:
        pytest.skip("reference extension did not travel")
    from pillarnet_lts_b200.compat import iou3d_nms_cuda as ours
    rng = np.random.default_rng(11)
    b = torch.from_numpy(rand_boxes(rng, 1500, spread=15.0, clusters=40)).cuda()
    ka, kb = torch.zeros(1500, dtype=torch.int64), torch.zeros(1500, dtype=torch.int64)
    na = ours.nms_normal_gpu(b, ka, 0.3)
    nb = ref.nms_normal_gpu(b, kb, 0.3)
    assert na == nb and torch.equal(ka[:na], kb[:nb])
    # input checks behave like the reference's CHECK_INPUT: a CPU tensor is refused
    with pytest.raises(RuntimeError):
        ours.nms_gpu(b.cpu(), ka, 0.3)

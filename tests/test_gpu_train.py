"""GPU parity of the training path: scatter-max forward/backward vs the reference's own
pillar_cuda extension, point features vs the torch expression of pillar_utils.py:51-56, sparse-conv forward /
data gradient / weight gradient vs torch autograd on the dense-equivalent convolution, aligned BEV overlap vs
the reference's iou3d_nms_cuda, and one whole training step of the detector.

Tolerances: fp32 (SIMT) 1e-4 relative to max|ref|; bf16 tensor-core kernels against the fp32 kernels fed the
same bf16-rounded operands: 5e-3 rel-to-max for dW (fp32 accumulation over thousands of rows, split-K order),
1e-2 for activations / data gradients (one bf16 rounding of the result)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import batch_points, rand_boxes, rand_points, ref_ext
from tests.test_gpu_rulebook_conv import _random_sites, _table_from_sites

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(1e-12, b.float().abs().max().item())


def test_point_features_vs_torch_expression():
    from pillarnet_lts_b200 import ops
    rng = np.random.default_rng(3)
    pts = torch.from_numpy(rand_points(rng, 5000, -53.9, 53.9)).cuda()
    ps, x0, y0 = 0.075, -54.0, -54.0
    xoff, yoff = ps / 2.0 + x0, ps / 2.0 + y0
    got = ops.point_features(pts, x0, y0, ps, xoff, yoff)
    # dynamic_pillar_encoder.py:34-35 + pillar_utils.py:51-56 with torch CUDA ops
    cx = torch.floor((pts[:, 0] - x0) / ps).int()
    cy = torch.floor((pts[:, 1] - y0) / ps).int()
    ctr = torch.stack([cx, cy], 1).float() * ps + torch.tensor([xoff, yoff], device="cuda")
    want = torch.cat([pts[:, :2] - ctr, pts], 1)
    assert torch.equal(got, want)


def test_scatter_max_forward_backward_vs_reference_extension():
    pillar = ref_ext("pillar_cuda")
    if pillar is None:
        pytest.skip("oracle/_ref/pillar_cuda not built")
    from pillarnet_lts_b200.autograd import scatter_max
    g = torch.Generator(device="cuda").manual_seed(9)
    L, C, M = 20000, 32, 1500
    src = torch.randn(L, C, device="cuda", generator=g)           # negatives too: floor at 0
    index = torch.randint(0, M - 7, (L,), device="cuda", generator=g, dtype=torch.int32)   # 7 empty pillars
    arg_ref = torch.full((M, C), -1, dtype=torch.int32, device="cuda")
    out_ref = torch.zeros(M, C, device="cuda")
    pillar.scatter_max_wrapper(index, src, arg_ref, out_ref)
    s = src.clone().requires_grad_(True)
    out = scatter_max(s, index, M)
    assert torch.equal(out, out_ref)                               # bit-exact forward
    go = torch.randn(M, C, device="cuda", generator=g)
    out.backward(go)
    grad_ref = torch.zeros(L, C, device="cuda")
    pillar.scatter_max_grad_wrapper(go, arg_ref, grad_ref)
    # The reference routes to ANY point within 1e-5 of the max (last writer wins); ours to the exact, lowest-index
    # argmax.  With continuous random values the two differ only at near-ties: allow a handful of entries.
    assert (s.grad != grad_ref).sum().item() <= 40
    assert torch.equal(s.grad.sum(0), grad_ref.sum(0)) or _rel(s.grad.sum(0), grad_ref.sum(0)) < 1e-5
    # own invariants: every positive pillar max is routed to a point of that pillar holding exactly that value
    routed = s.grad != 0
    p_idx = torch.nonzero(routed)
    assert torch.equal(out[index[p_idx[:, 0]].long(), p_idx[:, 1]], src[p_idx[:, 0], p_idx[:, 1]])


def _dense_equiv(x_rows, coords, B, H, W, weight, bias, stride, out_coords):
    """torch reference of a sparse conv: densify, F.conv2d, gather the active output sites"""
    c = coords.long()
    dense = x_rows.new_zeros(B, H, W, x_rows.shape[1])
    dense = dense.index_put((c[:, 0], c[:, 1], c[:, 2]), x_rows).permute(0, 3, 1, 2)
    y = F.conv2d(dense, weight.permute(0, 3, 1, 2), bias, stride=stride, padding=1)
    oc = out_coords.long()
    return y.permute(0, 2, 3, 1)[oc[:, 0], oc[:, 1], oc[:, 2]]


@pytest.mark.parametrize("precision,cin,cout,strided", [
    ("fp32", 32, 32, False), ("fp32", 16, 48, True), ("fp32", 64, 128, True),
    ("bf16", 32, 32, False), ("bf16", 64, 64, False), ("bf16", 64, 128, True), ("bf16", 128, 128, False),
    ("bf16", 128, 256, True), ("bf16", 256, 256, False)])
def test_sparse_conv_autograd_vs_torch_dense_equivalent(precision, cin, cout, strided):
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import train
    from pillarnet_lts_b200.autograd import sparse_conv
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    P.set_precision(precision)
    try:
        rng = np.random.default_rng(cin + cout)
        B, H, W, n = 2, 48, 40, 1500
        table = train._exact(_table_from_sites(_random_sites(rng, B, H, W, n), B, H, W))
        g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout)
        x = torch.randn(n, cin, device="cuda", generator=g)
        w = torch.randn(cout, 3, 3, cin, device="cuda", generator=g) * (1.0 / (3.0 * cin ** 0.5))
        b = torch.randn(cout, device="cuda", generator=g)
        if precision == "bf16":                                      # identical operands on both sides
            x, w = x.bfloat16().float(), w.bfloat16().float()
        if strided:
            out_table, rb = train.down_rulebook(table)
        else:
            out_table, rb = table, train.subm_rulebook(table)
        dy = torch.randn(out_table.cap, cout, device="cuda", generator=g)
        if precision == "bf16":
            dy = dy.bfloat16().float()
        xa, wa, ba = (t.clone().requires_grad_(True) for t in (x, w, b))
        y = sparse_conv(xa.to(P.config.act_dtype()), wa, ba, rb)
        y.backward(dy.to(y.dtype))
        xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
        yr = _dense_equiv(xr, table.coords, B, H, W, wr, br, 2 if strided else 1, out_table.coords)
        yr.backward(dy)
        torch.cuda.synchronize()
        tol_act, tol_w = (1e-4, 1e-4) if precision == "fp32" else (1e-2, 5e-3)
        assert _rel(y, yr) <= tol_act
        assert _rel(xa.grad, xr.grad) <= tol_act
        assert _rel(wa.grad, wr.grad) <= tol_w
        assert _rel(ba.grad, br.grad) <= tol_w
    finally:
        P.set_precision("bf16")


def test_rulebook_transpose_is_the_inverse_relation():
    from pillarnet_lts_b200 import ops, train
    rng = np.random.default_rng(5)
    B, H, W, n = 2, 33, 31, 700
    table = train._exact(_table_from_sites(_random_sites(rng, B, H, W, n), B, H, W))
    out_table, rb = train.down_rulebook(table)
    nbr, nbr_t = rb.nbr.cpu().numpy(), rb.nbr_t.cpu().numpy()
    want = np.full((n, 9), -1, np.int32)
    for o in range(nbr.shape[0]):
        for t in range(9):
            if nbr[o, t] >= 0:
                assert want[nbr[o, t], t] == -1          # (input, tap) pairs are unique
                want[nbr[o, t], t] = o
    assert np.array_equal(nbr_t, want)
    sub = train.subm_rulebook(table)
    assert np.array_equal(sub.nbr_t.cpu().numpy(), ops.rulebook_transpose(sub.nbr, n).cpu().numpy())


def test_aligned_overlap_vs_reference_extension():
    """IouLoss target (training only).  Tolerance 1e-5 rel-to-max: nvcc's FMA contraction of the polygon-area
    arithmetic differs by instantiation context (measured: 25 % of pairs differ in the last bit, max 1.9e-6), unlike
    the thresholded NMS / IoU kernels, which are bit-exact (tests/test_gpu_nms.py)."""
    iou3d = ref_ext("iou3d_nms_cuda")
    if iou3d is None:
        pytest.skip("oracle/_ref/iou3d_nms_cuda not built")
    from pillarnet_lts_b200 import ops
    rng = np.random.default_rng(8)
    a = rand_boxes(rng, 4000, spread=5.0)
    b = a + rng.normal(0, 0.3, a.shape).astype(np.float32)
    b[:500] = rand_boxes(rng, 500, spread=5.0)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    want = torch.zeros(len(a), 1, device="cuda")
    iou3d.boxes_aligned_overlap_bev_gpu(ta, tb, want)
    got = ops.boxes_aligned_overlap_bev(ta, tb)
    want = want.view(-1)
    # degenerate pairs (parallel edges, D == 0 in the line-intersection fallback) give NaN in both
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    g, w = torch.nan_to_num(got), torch.nan_to_num(want)
    assert (g - w).abs().max().item() <= 1e-5 * w.abs().max().item()
    assert (got > 0).sum().item() > 2000


def test_detector_training_step_runs_and_reaches_every_parameter():
    """PillarNet-18 (small grid) forward + loss + backward in bf16: finite loss, finite gradients everywhere."""
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import configs, synth, train
    torch.manual_seed(0)
    cfg = configs.get("nusc18")
    cfg["model"]["reader"]["pc_range"] = [-24.0, -24.0, -5.0, 24.0, 24.0, 3.0]
    cfg["model"]["bbox_head"]["point_cloud_range"] = [-24.0, -24.0, -5.0, 24.0, 24.0, 3.0]
    model = P.build_detector(cfg["model"], train_cfg=cfg["train_cfg"], test_cfg=cfg["test_cfg"]).cuda().train()
    rng = np.random.default_rng(1)
    frames = [rand_points(rng, 30000, -23.9, 23.9) for _ in range(2)]
    pts, off = batch_points(frames)
    example = {"points_batched": (pts, off), "points": None, "metadata": [None, None]}
    example.update(train.synthetic_targets(model.bbox_head, 2, 640, 640, rng, max_objs=50))
    losses = model(example, return_loss=True)
    loss = sum(l.sum() for l in losses["loss"])
    assert torch.isfinite(loss)
    loss.backward()
    missing = [n for n, p in model.named_parameters() if p.grad is None]
    assert not missing, missing
    assert all(torch.isfinite(p.grad).all() for p in model.parameters())
    # the gradient reaches the reader's Linear through the scatter-max backward and the sparse-conv dgrads
    assert model.reader.pfn_layers.shared_mlps[0].weight.grad.abs().max().item() > 0


def test_training_reduces_the_loss_on_a_fixed_batch():
    """five AdamW steps on one synthetic batch (bf16 kernels): the loss goes down and stays finite"""
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import configs, train
    torch.manual_seed(1)
    cfg = configs.get("nusc18")
    cfg["model"]["reader"]["pc_range"] = [-24.0, -24.0, -5.0, 24.0, 24.0, 3.0]
    cfg["model"]["bbox_head"]["point_cloud_range"] = [-24.0, -24.0, -5.0, 24.0, 24.0, 3.0]
    model = P.build_detector(cfg["model"], train_cfg=cfg["train_cfg"], test_cfg=cfg["test_cfg"]).cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
    rng = np.random.default_rng(2)
    pts, off = batch_points([rand_points(rng, 20000, -23.9, 23.9) for _ in range(2)])
    example = {"points_batched": (pts, off), "points": None, "metadata": [None, None]}
    example.update(train.synthetic_targets(model.bbox_head, 2, 640, 640, rng, max_objs=40))
    losses = [float(train.train_step(model, example, opt)) for _ in range(6)]
    assert all(np.isfinite(losses))
    assert losses[-1] < 0.9 * losses[0], losses


def test_assign_label_gpu_vs_reference_golden(golden_dir):
    """labels.AssignLabel (pn_assign_labels) on a 3-frame batch vs the reference's AssignLabel pipeline stage:
    ind / mask / cat / gt_box exact, heat-map within one fp32 ulp (double exp, rounded once), anno_box 1e-6."""
    import os
    from pillarnet_lts_b200.labels import AssignLabel
    g = np.load(os.path.join(golden_dir, "assign_label.npz"))
    tasks = [dict(stride=8, class_names=["car"]), dict(stride=4, class_names=["ped", "cone"])]
    al = AssignLabel(dict(target_assigner=dict(tasks=tasks), gaussian_overlap=0.1, max_objs=80, min_radius=2,
                          pc_range=[-24.0, -24.0, -5.0, 24.0, 24.0, 3.0], pillar_size=0.075))
    boxes = [torch.from_numpy(g[f"f{f}_boxes"]).cuda() for f in range(3)]
    cls = [torch.from_numpy(g[f"f{f}_cls"]).cuda() for f in range(3)]
    out = al(boxes, cls)
    torch.cuda.synchronize()
    for t in range(2):
        for f in range(3):
            for k in ("ind", "mask", "cat"):
                assert np.array_equal(out[k][t][f].cpu().numpy(), g[f"f{f}_t{t}_{k}"]), (f, t, k)
            assert np.array_equal(out["gt_box"][t][f].cpu().numpy(), g[f"f{f}_t{t}_gt_box"])
            hm, want = out["hm"][t][f].cpu().numpy(), g[f"f{f}_t{t}_hm"]
            assert hm.shape == want.shape
            assert np.array_equal(hm > 0, want > 0)
            np.testing.assert_allclose(hm, want, rtol=1.2e-7, atol=0)
            np.testing.assert_allclose(out["anno_box"][t][f].cpu().numpy(), g[f"f{f}_t{t}_anno_box"], rtol=2e-6, atol=2e-7)
    # the targets plug into the loss
    assert out["hm"][1].shape == (3, 160, 160, 2) and out["ind"][0].dtype == torch.int64


# ---- sync-free (CUDA-graph) training path: fused batch-stat BN kernels, static shapes, TrainEngine --------------------
def _small_train_setup(seed, precision, n_pts=20000, max_objs=40):
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import configs, train
    P.set_precision(precision)
    torch.manual_seed(seed)
    cfg = configs.get("nusc18")
    cfg["model"]["reader"]["pc_range"] = [-24.0, -24.0, -5.0, 24.0, 24.0, 3.0]
    cfg["model"]["bbox_head"]["point_cloud_range"] = [-24.0, -24.0, -5.0, 24.0, 24.0, 3.0]
    cfg["model"]["neck"]["layer_nums"] = [1, 1]
    model = P.build_detector(cfg["model"], train_cfg=cfg["train_cfg"], test_cfg=cfg["test_cfg"]).cuda().train()
    rng = np.random.default_rng(seed)
    batches = []
    for _ in range(3):
        pts, off = batch_points([rand_points(rng, n_pts, -26.0, 26.0) for _ in range(2)])   # some points out of range
        ex = {"points_batched": (pts, off), "points": None, "metadata": [None, None]}
        ex.update(train.synthetic_targets(model.bbox_head, 2, 640, 640, rng, max_objs=max_objs))
        batches.append(ex)
    return model, cfg, batches


def _grads(model):
    return {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}


def _rel(a, b):
    return (a - b).abs().max().item() / max(1e-6, b.abs().max().item())


def _pre_bn_biases(model):
    """conv biases in front of a batch-statistics BN: their gradient is mathematically zero (BN removes the mean), what
    either path computes for them is rounding noise"""
    names = set()
    for mn, m in model.named_modules():
        kids = list(m.children()) if isinstance(m, torch.nn.Sequential) else []
        for a, b in zip(kids, kids[1:]):
            if isinstance(b, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)) and getattr(a, "bias", None) is not None:
                names.add(f"{mn}.{kids.index(a)}.bias" if mn else f"{kids.index(a)}.bias")
    return names


def _grad_errors(got, want, skip=()):
    """(relative L2 error of the concatenated gradient vector, worst per-tensor max-abs error relative to the tensor's
    largest gradient over the weight tensors with >= 512 elements)"""
    keys = [k for k in want if k not in skip]
    g = torch.cat([got[k].reshape(-1) for k in keys])
    w = torch.cat([want[k].reshape(-1) for k in keys])
    l2 = ((g - w).norm() / w.norm()).item()
    worst = max((_rel(got[k], want[k]), k) for k in keys if want[k].numel() >= 512)
    return l2, worst


def test_bn_train_kernels_vs_torch_batch_norm():
    """pn_bn_stats / finalize / apply / bwd_stats / bwd_apply vs nn.BatchNorm1d (train) + add + ReLU under autograd on
    the live rows; rows beyond the device-resident count do not enter the statistics."""
    from pillarnet_lts_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    cap, n, C = 5000, 3777, 64
    num = torch.tensor([n], dtype=torch.int32, device="cuda")
    x = torch.randn(cap, C, device="cuda", generator=g) * 2 + 0.5
    res = torch.randn(cap, C, device="cuda", generator=g)
    dy = torch.randn(cap, C, device="cuda", generator=g)
    bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    xr, rr = x[:n].clone().requires_grad_(True), res[:n].clone().requires_grad_(True)
    want = torch.relu(bn(xr) + rr)
    want.backward(dy[:n])
    y, mean, rstd = ops.bn_train_forward(x, num, bn.weight.detach(), bn.bias.detach(), rm, rv, bn.eps, bn.momentum, res, True)
    dx, dres, dgamma, dbeta = ops.bn_train_backward(dy, y, x, mean, rstd, bn.weight.detach(), True, num, True)
    torch.cuda.synchronize()
    assert _rel(y[:n], want.detach()) <= 1e-5
    assert _rel(dx[:n], xr.grad) <= 1e-4 and _rel(dres[:n], rr.grad) <= 1e-6
    assert _rel(dgamma, bn.weight.grad) <= 1e-4 and _rel(dbeta, bn.bias.grad) <= 1e-4
    assert _rel(rm, bn.running_mean) <= 1e-5 and _rel(rv, bn.running_var) <= 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_dense_bn_relu_node_vs_torch_batch_norm_2d(dtype, tol):
    """autograd.DenseBNFunction (BatchNorm2d in train mode + ReLU on a channels-last map, library BN kernels) vs
    nn.BatchNorm2d + ReLU under torch autograd: outputs, input / weight / bias gradients, running statistics and
    num_batches_tracked; a Sequential with nested blocks through train.run_dense_seq."""
    import copy
    from pillarnet_lts_b200 import train
    from pillarnet_lts_b200.autograd import DenseBNFunction
    g = torch.Generator(device="cuda").manual_seed(12)
    B, C, H, W = 2, 64, 23, 31
    x = (torch.randn(B, C, H, W, device="cuda", generator=g) * 2 + 0.3).to(dtype).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, C, H, W, device="cuda", generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(C, eps=1e-3, momentum=0.01).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    bn2 = copy.deepcopy(bn)
    xr = x.clone().requires_grad_(True)
    want = torch.relu(bn(xr.float())).to(dtype)
    want.backward(dy)
    xg = x.clone().requires_grad_(True)
    got = DenseBNFunction.apply(xg, bn2.weight, bn2.bias, bn2, True)
    got.backward(dy)
    torch.cuda.synchronize()
    assert _rel(got.float(), want.float()) <= tol
    assert _rel(xg.grad.float(), xr.grad.float()) <= 5 * tol
    assert _rel(bn2.weight.grad, bn.weight.grad) <= 5 * tol and _rel(bn2.bias.grad, bn.bias.grad) <= 5 * tol
    assert _rel(bn2.running_mean, bn.running_mean) <= 1e-4 and _rel(bn2.running_var, bn.running_var) <= 1e-3
    assert int(bn2.num_batches_tracked) == int(bn.num_batches_tracked) == 1
    # a nested Sequential (conv, BN, ReLU, block(conv, BN, ReLU)): the fused nodes replace the BN + ReLU pairs only in the
    # static path, and both paths agree
    torch.manual_seed(3)
    seq = torch.nn.Sequential(torch.nn.Conv2d(C, 32, 3, padding=1, bias=True), torch.nn.BatchNorm2d(32, eps=1e-3), torch.nn.ReLU(),
                              torch.nn.Sequential(torch.nn.Conv2d(32, 32, 3, padding=1, bias=False),
                                                  torch.nn.BatchNorm2d(32, eps=1e-3, momentum=0.01), torch.nn.ReLU())).cuda().train()
    seq = seq.to(memory_format=torch.channels_last)
    xin = x.float()
    with torch.no_grad():
        seq[0].bias.normal_(0, 0.5)
    seq2 = copy.deepcopy(seq)
    ref = seq(xin)
    ref.sum().backward()
    try:
        train.set_static(True)
        out = train.run_dense_seq(seq2, xin)
    finally:
        train.set_static(False)
    out.sum().backward()
    assert _rel(out, ref) <= 1e-4
    # the conv bias in front of the batch norm is not added (BN removes it); the running mean still tracks it
    assert _rel(seq2[1].running_mean, seq[1].running_mean) <= 1e-4 and _rel(seq2[1].running_var, seq[1].running_var) <= 1e-3
    assert float(seq2[0].bias.grad.abs().max()) == 0.0 and float(seq[0].bias.grad.abs().max()) <= 1e-3
    assert _rel(seq2[0].weight.grad, seq[0].weight.grad) <= 1e-3


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-2), ("bf16", 1e-1)])
def test_static_training_path_matches_dynamic_path(precision, tol):
    """same parameters, same batch: the sync-free path (capacity-sized rows, fused BN kernels, masked PFN) gives the head
    maps and the parameter gradients of the exactly sized path (torch BatchNorm1d on compacted rows).  The probe loss is
    LINEAR in the head maps (fixed random weights): CenterHead's L1 terms have sign discontinuities, which turn a
    rounding-level difference of the predictions into O(1/num_pos) jumps of the gradient and make a per-step
    comparison of two correct bf16 paths meaningless (the real loss is compared in fp32 against the dense-equivalent
    autograd model below)."""
    from pillarnet_lts_b200 import train
    model, cfg, batches = _small_train_setup(5, precision)
    ex = batches[0]
    out = {}
    probe = None
    try:
        for mode in (False, True):
            train.set_static(mode)
            model.zero_grad(set_to_none=True)
            bev, _ = model.extract_feat(dict(points_batched=ex["points_batched"], points=None))
            preds = model.bbox_head(bev)
            if probe is None:
                g = torch.Generator(device="cuda").manual_seed(9)
                probe = [{k: torch.randn(v.shape, device="cuda", generator=g) for k, v in p.items()} for p in preds]
            loss = sum((p[k].float() * w[k]).sum() for p, w in zip(preds, probe) for k in p) / 1000.0
            loss.backward()
            out[mode] = (float(loss.detach()), _grads(model), [{k: v.detach().float().clone() for k, v in p.items()} for p in preds])
    finally:
        train.set_static(False)
    if precision == "fp32":
        # the probe loss is a sum of ~1e6 random-signed terms with heavy cancellation: its VALUE is only comparable where
        # the maps agree to 1e-3 (fp32); in bf16 the maps themselves are compared below (rel-to-max)
        assert abs(out[True][0] - out[False][0]) <= tol * max(1.0, abs(out[False][0]))
    assert set(out[True][1]) == set(out[False][1])
    # Tolerances: each operator is pinned tightly on its own (BN kernels 1e-5 / 1e-4 above, conv fwd / dgrad / wgrad
    # 1e-4 in fp32); through ~25 batch-norm layers in TRAIN mode a 1e-6 difference in one conv output is amplified
    # (rstd up to 1/sqrt(eps) = 31 per layer on low-variance channels, ReLU masks flipping near zero): measured fp32
    # head maps agree to ~1e-3 and the full gradient vector to 6e-3 relative L2 — between ANY two implementations,
    # e.g. this path and cuDNN through the dense-equivalent model below.  A wiring bug gives O(1).
    l2, worst = _grad_errors(out[True][1], out[False][1], skip=_pre_bn_biases(model))
    fwd = max(_rel(a[k], b[k]) for a, b in zip(out[True][2], out[False][2]) for k in a)
    if precision == "fp32":
        assert fwd <= 5e-3 and l2 <= tol * 2 and worst[0] <= 0.3, (fwd, l2, worst)
    else:
        # bf16: the same amplification acts on 4e-3 roundings — head maps within 1e-1 (measured 6e-2), while the gradient of the FIRST
        # layers (backward through every BN of the net) differs by tens of percent between two correct bf16 paths
        # (measured l2 0.49).  Only the wiring is asserted here (the gradient vectors point the same way); the bf16
        # kernels are pinned per operator in the tests above.
        g = torch.cat([v.reshape(-1) for k, v in out[True][1].items()])
        w = torch.cat([out[False][1][k].reshape(-1) for k in out[True][1]])
        cos = float((g * w).sum() / (g.norm() * w.norm()))
        assert fwd <= 1e-1 and cos >= 0.7, (fwd, cos, l2)


def test_whole_step_gradients_vs_dense_equivalent_autograd():
    """gradients of EVERY backbone / neck / head parameter from the library's training path (fp32
    mode: gather conv fwd / dgrad / wgrad kernels + fused BN kernels) against torch autograd through the dense-equivalent
    model (oracle/dense_train.py: F.conv2d + batch-norm over the active rows), same reader output, same loss."""
    from oracle import dense_train
    from pillarnet_lts_b200 import train
    model, cfg, batches = _small_train_setup(6, "fp32", n_pts=12000, max_objs=30)
    ex = batches[0]
    try:
        train.set_static(True)
        model.zero_grad(set_to_none=True)
        sp = model.reader(dict(points_batched=ex["points_batched"]))
        feat_leaf = sp.feat.detach().clone().requires_grad_(True)
        sp.feat = feat_leaf
        feats = model.backbone(sp)
        bev = model.neck(feats)
        preds = model.bbox_head(bev)
        loss = sum(l.sum() for l in model.bbox_head.loss(ex, preds, cfg["train_cfg"])["loss"])
        loss.backward()
        got = _grads(model)
        got_feat = feat_leaf.grad.clone()
        n = sp.table.count()
    finally:
        train.set_static(False)
    model.zero_grad(set_to_none=True)
    leaf = feat_leaf.detach()[:n].clone().requires_grad_(True)
    want_loss = dense_train.loss_dense_equivalent(model, leaf, sp.table.coords[:n], 2, sp.table.H, sp.table.W, ex,
                                                  cfg["train_cfg"])
    want_loss.backward()
    want = _grads(model)
    assert abs(float(loss) - float(want_loss)) <= 1e-3 * abs(float(want_loss))
    assert _rel(got_feat[:n], leaf.grad) <= 2e-2
    skip = _pre_bn_biases(model) | {k for k in want if k.startswith("reader.")}
    assert len([k for k in want if k not in skip]) > 100
    l2, worst = _grad_errors(got, want, skip=skip)
    assert l2 <= 2e-2 and worst[0] <= 0.2, (l2, worst)


def test_train_engine_graph_replay_matches_eager_steps():
    """TrainEngine (two CUDA graphs, fixed input buffers) against eager steps of the same static path from the same
    initial state: the loss curves agree (fp32 atomics in the statistics kernels reorder sums, hence not bit-equal) and
    the loss goes down."""
    import copy
    from pillarnet_lts_b200 import train
    model, cfg, batches = _small_train_setup(7, "bf16")
    model2 = copy.deepcopy(model)
    cap = max(b["points_batched"][0].shape[0] for b in batches) + 1000
    try:
        opt = torch.optim.AdamW(model.parameters(), lr=2e-4, capturable=True)
        eng = train.TrainEngine(model, opt, 2, cap, batches[0]).prepare(warmup=2)
        got = []
        for i in range(6):
            loss = eng.step(batches[i % 3])
            eng.stream.synchronize()            # the step runs on the engine's stream
            got.append(float(loss))
        train.set_static(True)
        opt2 = torch.optim.AdamW(model2.parameters(), lr=2e-4, capturable=True)
        eng2 = train.TrainEngine(model2, opt2, 2, cap, batches[0], use_graph=False).prepare(warmup=2)
        want = []
        for i in range(6):
            loss = eng2.step(batches[i % 3])
            eng2.stream.synchronize()
            want.append(float(loss))
    finally:
        train.set_static(False)
    assert all(np.isfinite(got)) and all(np.isfinite(want))
    # two runs of the same bf16 train-mode model drift apart step by step (batch-statistics BN amplifies the reordered
    # fp32 atomics and every later optimiser step inherits it): 2 % at the first step, 3-5.3 % by the sixth measured
    assert max(abs(a - b) / abs(b) for a, b in zip(got, want)) <= 1e-1, (got, want)
    assert abs(got[0] - want[0]) / abs(want[0]) <= 5e-2, (got, want)
    assert got[-1] < got[0]

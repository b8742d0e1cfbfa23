"""TEST INFRASTRUCTURE ONLY — CPU (torch) restatement of the Pillar R-CNN second stage, inference path.

Only tests/ may import this file; the product path never does.  Each function cites what it restates:
  grid_points        det3d/core/bbox/box_torch_ops.py:220-251 (center_to_grid_box2d, get_dense_roi_grid_points) +
                     rotation_2d :159-172
  bilinear           det3d/core/utils/center_utils.py:91-120 (bilinear_interpolate_torch)
  deconv_ks / block_sparse_conv / bev_fusion
                     det3d/models/second_stage/bev_interpolation.py:39-83,125-159 (BEVFeature) and :186-231,273-308
                     (BEVStrideFeature): ConvTranspose2d(k = s, stride = s) + BN + ReLU, spconv SparseConv2d(k = s,
                     stride = s) + BN1d + ReLU restated densely (output cell active iff its s x s input block holds an
                     active cell; bias only on active outputs), concat, 3x3 fusion conv + BN + ReLU
  fc_stack           det3d/models/roi_heads/roi_head_template.py:23-39, roi_mix_head.py:36-59,101-105
  refine             roi_head_template.py:189-219 (generate_predicted_boxes) + detectors/pillar_rcnn.py:141-170
Pinned against tests/golden/second_stage.npz (produced by executing the reference's own modules,
tests/golden/make_golden.py::gen_second_stage) in tests/test_second_stage.py; the sparse lateral conv is unpinned
(spconv is not in /root/reference) and follows spconv's published semantics like the backbone oracle.
"""
import torch
import torch.nn.functional as F


def grid_points(rois, G):
    """rois (n, >=7) -> (n, G*G, 2); point p = i*G + j with i the x index (torch.nonzero order of a GxG ones map)"""
    n = rois.shape[0]
    idx = torch.ones(G, G).nonzero().float().unsqueeze(0).repeat(n, 1, 1)          # (n, G*G, 2) [x_idx, y_idx]
    dims = rois[:, 3:5].view(n, 1, 2)
    pts = (idx + 0.5) / torch.tensor([G, G], dtype=torch.float32) * dims - dims / 2
    ang = rois[:, -1]
    s, c = torch.sin(ang), torch.cos(ang)
    x = pts[..., 0] * c[:, None] + pts[..., 1] * s[:, None]
    y = -pts[..., 0] * s[:, None] + pts[..., 1] * c[:, None]
    return torch.stack([x, y], -1) + rois[:, :2].view(n, 1, 2)


def bilinear(im, x, y):
    """im (H, W, C); x, y (n) map coordinates -> (n, C) with the reference's clamped corners"""
    x0 = torch.floor(x).long()
    y0 = torch.floor(y).long()
    x1, y1 = x0 + 1, y0 + 1
    x0, x1 = x0.clamp(0, im.shape[1] - 1), x1.clamp(0, im.shape[1] - 1)
    y0, y1 = y0.clamp(0, im.shape[0] - 1), y1.clamp(0, im.shape[0] - 1)
    wa = (x1.float() - x) * (y1.float() - y)
    wb = (x1.float() - x) * (y - y0.float())
    wc = (x - x0.float()) * (y1.float() - y)
    wd = (x - x0.float()) * (y - y0.float())
    return im[y0, x0] * wa[:, None] + im[y1, x0] * wb[:, None] + im[y0, x1] * wc[:, None] + im[y1, x1] * wd[:, None]


def _bn(x, w, b, mean, var, eps, dim):
    shape = [1] * x.dim()
    shape[dim] = -1
    return (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + eps) * w.view(shape) + b.view(shape)


def deconv_ks(x, weight, bn, eps=1e-3):
    """x (B, Cin, H, W); weight (Cin, Cout, s, s); bn = (w, b, mean, var)"""
    s = weight.shape[2]
    return F.relu(_bn(F.conv_transpose2d(x, weight, stride=s), *bn, eps, 1))


def block_sparse_conv(x, active, weight, bias, bn, eps=1e-3):
    """dense restatement of SparseConv2d(k = s, stride = s, bias=True) + BN1d + ReLU followed by .dense():
    x (B, Cin, H, W) zero at inactive cells, active (B, H, W) bool, weight (Cout, s, s, Cin) spconv layout"""
    s = weight.shape[1]
    H, W = x.shape[2] // s * s, x.shape[3] // s * s
    y = F.conv2d(x[:, :, :H, :W], weight.permute(0, 3, 1, 2), bias, stride=s)
    act = F.max_pool2d(active[:, None, :H, :W].float(), s, stride=s) > 0
    return F.relu(_bn(y, *bn, eps, 1)) * act


def roi_pool(fused, rois, G, x0, y0, cell):
    """fused (B, C, H, W), rois (B, N, >=7) -> (features (B, N, G*G, C), points (B, N, G*G, 2))"""
    B, N = rois.shape[:2]
    pts = grid_points(rois.reshape(B * N, -1), G).view(B, N, G * G, 2)
    xs = (pts[..., 0] - x0) / cell
    ys = (pts[..., 1] - y0) / cell
    out = []
    for b in range(B):
        im = fused[b].permute(1, 2, 0)
        out.append(bilinear(im, xs[b].reshape(-1), ys[b].reshape(-1)).view(N, G * G, -1))
    return torch.stack(out), pts


def fc_stack(x, layers):
    """x (n, C); layers: list of dicts {weight (out, in), bias | None, bn (w, b, mean, var) | None, relu}"""
    for l in layers:
        x = x @ l["weight"].t()
        if l.get("bias") is not None:
            x = x + l["bias"]
        if l.get("bn") is not None:
            x = _bn(x, *l["bn"], 1e-3, 1)
        if l.get("relu"):
            x = F.relu(x)
    return x


def refine(rois, reg, cls, roi_scores, roi_labels):
    """-> boxes (B, N, code), scores (B, N), valid (B, N)"""
    B, N, _ = rois.shape
    code = reg.shape[-1]
    reg = reg.view(B, N, code)
    local = rois[..., :code].clone()
    local[..., 0:3] = 0
    p = (reg + local).view(-1, code)
    ang = rois[..., 6].reshape(-1)
    c, s = torch.cos(ang), torch.sin(ang)
    x = p[:, 0] * c + p[:, 1] * s
    y = -p[:, 0] * s + p[:, 1] * c
    boxes = torch.cat([torch.stack([x, y, p[:, 2]], 1) + rois[..., 0:3].reshape(-1, 3), p[:, 3:]], 1).view(B, N, code)
    scores = torch.sqrt(torch.sigmoid(cls.view(B, N)) * roi_scores)
    valid = (roi_labels != 0) & (boxes[..., 3:6] > 0).all(-1)
    return boxes, scores, valid

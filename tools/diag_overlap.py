import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.gpu_util import rand_boxes, ref_ext
from pillarnet_lts_b200 import ops
iou3d = ref_ext("iou3d_nms_cuda")
rng = np.random.default_rng(8)
a = rand_boxes(rng, 1500, spread=5.0)
b = a + rng.normal(0, 0.3, a.shape).astype(np.float32)
ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
al = torch.zeros(len(a), 1, device="cuda"); iou3d.boxes_aligned_overlap_bev_gpu(ta, tb, al); al = al.view(-1)
full = torch.zeros(len(a), len(a), device="cuda"); iou3d.boxes_overlap_bev_gpu(ta, tb, full); dg = full.diagonal().contiguous()
mine = ops.boxes_aligned_overlap_bev(ta, tb)
iou_ref = torch.zeros(len(a), len(a), device="cuda"); iou3d.boxes_iou_bev_gpu(ta, tb, iou_ref)
iou_mine = ops.boxes_iou_bev(ta, tb)
def mm(x, y): 
    x, y = torch.nan_to_num(x), torch.nan_to_num(y)
    return int((x != y).sum())
print("ref aligned vs ref matrix diag:", mm(al, dg))
print("mine vs ref aligned:", mm(mine, al), " mine vs ref matrix diag:", mm(mine, dg))
print("iou mine vs ref (full matrix):", mm(iou_mine, iou_ref))

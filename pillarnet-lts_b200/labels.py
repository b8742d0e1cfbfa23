"""GPU training-target assignment behind the reference's `AssignLabel` pipeline stage.

Mirrors det3d/datasets/pipelines/preprocess.py:177-350 (same assigner cfg keys: target_assigner.tasks,
gaussian_overlap, max_objs, min_radius, pc_range, pillar_size) for a whole BATCH on the device: boxes are regrouped
per task in the reference's order (per class of the task, original order within a class, :203-231), headings are
wrapped to [-pi, pi) (:233-237) and every task's heat-map / ind / mask / cat / anno_box / gt_box come from one
pn_assign_labels launch.  The result has the layout the collated example has (`CenterHead.loss` input).
"""
import numpy as np
import torch

from . import ops


def limit_period(val, offset=0.5, period=np.pi):
    """det3d/core/bbox/box_np_ops.py: val - floor(val / period + offset) * period"""
    return val - torch.floor(val / period + offset) * period


class AssignLabel:
    def __init__(self, cfg):
        self.tasks = cfg["target_assigner"]["tasks"]
        self.gaussian_overlap = cfg["gaussian_overlap"]
        self.max_objs = cfg["max_objs"]
        self.min_radius = cfg["min_radius"]
        self.pc_range = [float(v) for v in cfg["pc_range"]]
        self.pillar_size = float(cfg["pillar_size"])

    def __call__(self, gt_boxes, gt_classes):
        """gt_boxes: list (per frame) of (n_i, 9|7) f32 CUDA tensors; gt_classes: list of (n_i,) int tensors with
        the global 1-based class ids of the reference (class j of task t has id 1 + sum(len(tasks[:t])) + j).
        Returns the example dict: hm, anno_box, ind, mask, cat, gt_box — lists over tasks of batched tensors."""
        pcr = np.array(self.pc_range, dtype=np.float32)
        ps = np.float32(self.pillar_size)
        grid = np.round((pcr[3:5] - pcr[:2]) / ps).astype(np.int64)       # (W, H)  preprocess.py:196-199
        B, M = len(gt_boxes), self.max_objs
        dev = gt_boxes[0].device
        D = gt_boxes[0].shape[1]
        out = {k: [] for k in ("hm", "anno_box", "ind", "mask", "cat", "gt_box")}
        flag = 0
        for task in self.tasks:
            names = task["class_names"]
            stride = int(task["stride"])
            boxes = torch.zeros(B, M, D, dtype=torch.float32, device=dev)
            cls = torch.zeros(B, M, dtype=torch.int32, device=dev)
            for b in range(B):
                gb, gc = gt_boxes[b].float(), gt_classes[b].to(torch.int64)
                parts_b, parts_c = [], []
                for j in range(len(names)):                     # per class of the task, original order inside
                    sel = torch.nonzero(gc == j + 1 + flag).squeeze(1)
                    parts_b.append(gb.index_select(0, sel))
                    parts_c.append(torch.full((sel.numel(),), j + 1, dtype=torch.int32, device=dev))
                tb, tc = torch.cat(parts_b), torch.cat(parts_c)
                tb = tb.clone()
                tb[:, -1] = limit_period(tb[:, -1], offset=0.5, period=np.pi * 2)
                n = min(tb.shape[0], M)
                boxes[b, :n], cls[b, :n] = tb[:n], tc[:n]
            W, H = int(grid[0] // stride), int(grid[1] // stride)
            cell = float(np.float32(ps * stride))
            r = ops.assign_labels_task(boxes, cls, len(names), H, W, self.pc_range[0], self.pc_range[1], cell,
                                       self.gaussian_overlap, self.min_radius)
            for k in out:
                out[k].append(r[k])
            flag += len(names)
        return out

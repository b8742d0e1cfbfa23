"""Multi-GPU plumbing for the inference path: frame sharding + the final detection gather.

Frames are independent (the batch index is only the leading rulebook coordinate), so the path shards
with no data-path collective.  Semantics follow the reference:
  * frame i of the dataset goes to rank i % world (DistributedSampler, det3d/datasets/loader/sampler.py:74-96)
  * detections are gathered once at the end (det3d/torchie/trainer/utils.py:114-154 pickles python
    objects through a ByteTensor all_gather; here it is one fixed-shape tensor all-gather, no pickle).
"""
import torch
import torch.distributed as dist


def shard_indices(n_total, rank, world):
    """indices of the frames rank `rank` owns (strided, like the reference's sampler)"""
    return list(range(rank, n_total, world))


def unshard_order(n_total, world):
    """position in the rank-major gathered list of each global frame index"""
    pos, k = {}, 0
    for r in range(world):
        for i in shard_indices(n_total, r, world):
            pos[i] = k
            k += 1
    return [pos[i] for i in range(n_total)]


def gather_detections(det_out, keep_count, group=None):
    """All ranks contribute (n_segs, post_cap, 11) f32 + (n_segs,) i32 of identical shape;
    returns ((world, n_segs, post_cap, 11), (world, n_segs)) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return det_out.unsqueeze(0), keep_count.unsqueeze(0)
    dets = [torch.empty_like(det_out) for _ in range(world)]
    cnts = [torch.empty_like(keep_count) for _ in range(world)]
    dist.all_gather(dets, det_out.contiguous(), group=group)
    dist.all_gather(cnts, keep_count.contiguous(), group=group)
    return torch.stack(dets), torch.stack(cnts)


def merge_gathered(dets, cnts, frames_per_rank, segs_per_frame, n_total):
    """(world, B*S, P, 11),(world, B*S) -> per global frame list of (S, P, 11)/(S,) in dataset order."""
    world = dets.shape[0]
    out = [None] * n_total
    for r in range(world):
        idx = shard_indices(n_total, r, world)
        for j, g in enumerate(idx[:frames_per_rank]):
            sl = slice(j * segs_per_frame, (j + 1) * segs_per_frame)
            out[g] = (dets[r, sl], cnts[r, sl])
    return out

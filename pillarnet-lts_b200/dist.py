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


# ---- training: data-parallel gradient averaging -----------------------------------

def allreduce_grads(params, coalesce=True, bucket_size_mb=-1, group=None):
    """det3d/core/utils/dist_utils.py:31-42: average `.grad` over the ranks (flat buckets when coalesce)."""
    grads = [p.grad.data for p in params if p.requires_grad and p.grad is not None]
    world = dist.get_world_size(group)
    if not coalesce:
        for g in grads:
            dist.all_reduce(g.div_(world), group=group)
        return
    limit = bucket_size_mb * 1024 * 1024 if bucket_size_mb > 0 else None
    buckets, cur, size = {}, [], 0
    if limit is None:
        for g in grads:                              # one bucket per dtype (dist_utils.py:12-28)
            buckets.setdefault(g.dtype, []).append(g)
        groups = list(buckets.values())
    else:
        groups = []
        for g in grads:
            cur.append(g)
            size += g.numel() * g.element_size()
            if size >= limit:
                groups.append(cur)
                cur, size = [], 0
        if cur:
            groups.append(cur)
    for bucket in groups:
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


class GradientAverager:
    """Overlapped data-parallel gradient averaging: what torch DDP does for the reference
    (det3d/torchie/apis/train.py:283-290), as explicit plumbing.  Parameters are packed (reverse registration
    order ~ backward order) into flat fp32 buckets whose slices ARE the `.grad` tensors; a bucket's
    all-reduce is launched asynchronously (NCCL stream) the moment its last gradient has been accumulated, so
    the exchange of the head/neck gradients overlaps the backward of the backbone.  `finish()` waits and divides.
    """

    def __init__(self, params, bucket_mb=25, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []          # (flat tensor, [params])
        limit = int(bucket_mb * 1024 * 1024)
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= limit:
                self._make_bucket(cur)
                cur, size = [], 0
        if cur:
            self._make_bucket(cur)
        self._pending = [0] * len(self.buckets)
        self._works = []
        self._hooks = []
        for bi, (_, ps) in enumerate(self.buckets):
            for p in ps:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
        self.reset()

    def _make_bucket(self, ps):
        n = sum(p.numel() for p in ps)
        flat = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
        off = 0
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)     # gradients accumulate straight into the bucket
            off += p.numel()
        self.buckets.append((flat, list(ps)))

    def _make_hook(self, bi):
        def hook(_p):
            self._pending[bi] -= 1
            if self._pending[bi] == 0 and self.world > 1:
                self._works.append(dist.all_reduce(self.buckets[bi][0], group=self.group, async_op=True))
        return hook

    def reset(self):
        """call before each backward (after zeroing): re-arms the per-bucket counters"""
        self._pending = [len(ps) for _, ps in self.buckets]
        self._works = []

    def zero_grad(self):
        for flat, _ in self.buckets:
            flat.zero_()
        self.reset()

    def finish(self):
        """waits for the launched all-reduces, reduces buckets whose hooks did not all fire (unused
        parameters), and turns sums into means"""
        for bi, n in enumerate(self._pending):
            if n != 0 and self.world > 1:
                self._works.append(dist.all_reduce(self.buckets[bi][0], group=self.group, async_op=True))
        for w in self._works:
            w.wait()
        if self.world > 1:
            for flat, _ in self.buckets:
                flat.div_(self.world)
        self._works = []

    def remove(self):
        for h in self._hooks:
            h.remove()

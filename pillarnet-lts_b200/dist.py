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
    all-reduce is launched asynchronously (NCCL stream) once its last gradient has been accumulated, so
    the exchange of the head/neck gradients overlaps the backward of the backbone.  `finish()` waits and divides.

    Collective order is FIXED: bucket k's all-reduce is issued only after buckets 0..k-1 have been issued, and
    `finish()` flushes whatever is left in index order — so every rank issues the same sequence of collectives
    over the same tensors even when the set of parameters that received a gradient differs between ranks (unused
    branches), which would otherwise dead-lock or mix up buckets.  Like DDP's constructor, `__init__` broadcasts
    rank 0's parameters (and, when `module` is given, its buffers: BN running statistics) so the replicas start
    identical whatever their seeds.  `finish()` re-attaches a `.grad` that was detached from its bucket (e.g. by
    `optimizer.zero_grad(set_to_none=True)` outside `train_step`), folding the stray gradient into the bucket.
    """

    def __init__(self, params, bucket_mb=25, group=None, module=None, broadcast=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        if broadcast and self.world > 1:
            src = dist.get_global_rank(group, 0) if group is not None else 0
            with torch.no_grad():
                tensors = [p.data for p in params]
                if module is not None:
                    tensors += [b.data for b in module.buffers()]
                for t in tensors:
                    dist.broadcast(t, src=src, group=group)
        self.buckets = []          # (flat tensor, [params])
        self._slices = []          # per bucket: [(param, offset)]
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
        self.defer = False         # True: hooks only count (captured backward); finish() issues every all-reduce
        self._next = 0             # buckets [0, _next) have had their all-reduce issued this step
        self._works = []
        self._hooks = []
        for bi, (_, ps) in enumerate(self.buckets):
            for p in ps:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
        self.reset()

    def _make_bucket(self, ps):
        n = sum(p.numel() for p in ps)
        flat = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
        off, sl = 0, []
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)     # gradients accumulate straight into the bucket
            sl.append((p, off))
            off += p.numel()
        self.buckets.append((flat, list(ps)))
        self._slices.append(sl)

    def _issue_ready(self):
        """issue, in index order, the all-reduce of every leading bucket whose gradients are complete"""
        if self.defer:
            return
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            if self.world > 1:
                self._works.append(dist.all_reduce(self.buckets[self._next][0], group=self.group, async_op=True))
            self._next += 1

    def _make_hook(self, bi):
        flat = self.buckets[bi][0]
        offs = {id(p): off for p, off in self._slices[bi]}

        def hook(p):
            g = p.grad
            lo = flat.data_ptr()
            if g is not None and not (lo <= g.data_ptr() < lo + flat.numel() * 4):
                # .grad was detached from the bucket (set_to_none / replaced): fold this step's gradient in and
                # re-attach BEFORE the bucket can be reduced, instead of all-reducing stale zeros
                off = offs[id(p)]
                view = flat[off:off + p.numel()].view_as(p)
                view.add_(g.to(view.dtype))
                p.grad = view
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._issue_ready()
        return hook

    def reset(self):
        """call before each backward (after zeroing): re-arms the per-bucket counters"""
        self._pending = [len(ps) for _, ps in self.buckets]
        self._next = 0
        self._works = []

    def zero_grad(self):
        self._reattach(fold=False)
        for flat, _ in self.buckets:
            flat.zero_()
        self.reset()

    def _reattach(self, fold=True):
        """every `.grad` must alias its bucket slice; a detached one (set_to_none / replaced tensor) is folded
        into the bucket (fold=True: its value is a gradient of this step) and re-attached"""
        for flat, sl in zip((b[0] for b in self.buckets), self._slices):
            base, end = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
            for p, off in sl:
                view = flat[off:off + p.numel()].view_as(p)
                g = p.grad
                if g is None or not (base <= g.data_ptr() < end):
                    if g is not None and fold:
                        view.add_(g.to(view.dtype))
                    p.grad = view

    def finish(self):
        """flushes, in index order, the buckets not yet issued (hooks that never fired: unused parameters, or a
        bucket waiting behind one of those), waits, and turns sums into means"""
        self._reattach(fold=True)
        if self.defer:
            self._next = 0         # a replayed (captured) backward does not run the hooks: reduce every bucket here
        for bi in range(self._next, len(self.buckets)):
            if self.world > 1:
                self._works.append(dist.all_reduce(self.buckets[bi][0], group=self.group, async_op=True))
        self._next = len(self.buckets)
        for w in self._works:
            w.wait()
        if self.world > 1:
            for flat, _ in self.buckets:
                flat.div_(self.world)
        self._works = []

    def remove(self):
        for h in self._hooks:
            h.remove()

"""CPU, world_size 2 over gloo: frame sharding + detection gather host logic (the N>1 inference path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pillarnet_lts_b200  # noqa: F401
    from pillarnet_lts_b200 import dist as pd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S, P = 3, 5
    mine = pd.shard_indices(n_total, rank, world)
    det = torch.zeros(len(mine) * S, P, 11)
    cnt = torch.zeros(len(mine) * S, dtype=torch.int32)
    for j, g in enumerate(mine):          # "detections" of global frame g: marked with g
        det[j * S:(j + 1) * S] = float(g)
        cnt[j * S:(j + 1) * S] = g % 4
    dets, cnts = pd.gather_detections(det, cnt)
    merged = pd.merge_gathered(dets, cnts, len(mine), S, n_total)
    ok = all(float(m[0].mean()) == float(g) and int(m[1][0]) == g % 4 for g, m in enumerate(merged))
    ret[rank] = ok and dets.shape == (world, len(mine) * S, P, 11)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    world, n_total = 2, 8
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_shard_indices_are_a_partition():
    import pillarnet_lts_b200  # noqa: F401
    from pillarnet_lts_b200 import dist as pd
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 64):
            parts = [pd.shard_indices(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            order = pd.unshard_order(n, world)
            assert sorted(order) == list(range(n))


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pillarnet_lts_b200  # noqa: F401
    from pillarnet_lts_b200 import dist as pd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                       # same init on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 3))
    unused = torch.nn.Parameter(torch.ones(5))           # never receives a gradient
    params = list(net.parameters()) + [unused]
    g = torch.Generator().manual_seed(100)
    data = torch.randn(world, 8, 6, generator=g)          # rank r trains on data[r]
    # reference result: mean over ranks of the per-rank gradients
    want = []
    for r in range(world):
        net.zero_grad()
        net(data[r]).pow(2).mean().backward()
        want.append([p.grad.clone() for p in net.parameters()])
    want = [sum(gs) / world for gs in zip(*want)]
    ok = True
    # (1) overlapped bucketed averager (tiny buckets: several all-reduces launched from backward hooks)
    avg = pd.GradientAverager(params, bucket_mb=0.0002)
    assert len(avg.buckets) > 2
    for _ in range(2):                                    # two steps: counters re-arm, buffers are reused
        avg.zero_grad()
        net(data[rank]).pow(2).mean().backward()
        avg.finish()
        ok = ok and all(torch.allclose(p.grad, w, atol=1e-6) for p, w in zip(net.parameters(), want))
        ok = ok and bool((unused.grad == 0).all())
    avg.remove()
    # (2) the reference's after-backward helper, coalesced and bucketed
    for kw in (dict(coalesce=True, bucket_size_mb=-1), dict(coalesce=True, bucket_size_mb=1), dict(coalesce=False)):
        for p in params:
            p.grad = None
        net(data[rank]).pow(2).mean().backward()
        pd.allreduce_grads(params, **kw)
        ok = ok and all(torch.allclose(p.grad, w, atol=1e-6) for p, w in zip(net.parameters(), want))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_averaging_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def _robust_worker(rank, world, port, ret):
    """rank-dependent sets of used parameters, differently seeded replicas, detached .grad tensors"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pillarnet_lts_b200  # noqa: F401
    from pillarnet_lts_b200 import dist as pd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(10 + rank)                              # replicas start DIFFERENT
    trunk = torch.nn.Linear(6, 8)
    heads = torch.nn.ModuleList([torch.nn.Linear(8, 2), torch.nn.Linear(8, 2)])
    bn = torch.nn.BatchNorm1d(8)
    bn.running_mean.fill_(float(rank + 1))
    net = torch.nn.ModuleList([trunk, heads, bn])
    params = list(net.parameters())
    avg = pd.GradientAverager(params, bucket_mb=0.0001, module=net)
    ok = len(avg.buckets) >= 3
    # (b) rank 0's parameters and buffers everywhere
    flat = torch.cat([p.data.reshape(-1) for p in params] + [bn.running_mean])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok = ok and all(torch.equal(g, gathered[0]) for g in gathered) and float(bn.running_mean[0]) == 1.0
    g = torch.Generator().manual_seed(5)
    data = torch.randn(world, 4, 6, generator=g)

    def loss_of(r):                                            # rank r only uses head r
        return heads[r](torch.relu(trunk(data[r]))).pow(2).mean()

    # expected result on a hook-free copy of the (broadcast) replica: mean over ranks of the per-rank gradients
    import copy
    t2, h2 = copy.deepcopy(trunk), copy.deepcopy(heads)
    p2 = list(t2.parameters()) + list(h2.parameters())
    want = []
    for r in range(world):
        for p in p2:
            p.grad = None
        h2[r](torch.relu(t2(data[r]))).pow(2).mean().backward()
        want.append([torch.zeros_like(p) if p.grad is None else p.grad.clone() for p in p2])
    want = [sum(gs) / world for gs in zip(*want)] + [torch.zeros(8), torch.zeros(8)]     # + the unused BN affine
    for step in range(3):
        avg.zero_grad()
        if step == 1:
            for p in params:
                p.grad = None                                  # (c) every .grad detached from its bucket
        elif step == 2:
            for p in trunk.parameters():
                p.grad = None                                  # an optimizer.zero_grad(set_to_none=True) in between
        loss_of(rank).backward()                               # (a) different hooks fire on different ranks
        avg.finish()
        ok = ok and all(p.grad is not None and torch.allclose(p.grad, w, atol=1e-6) for p, w in zip(params, want))
        lo = min(b[0].data_ptr() for b in avg.buckets)
        ok = ok and all(p.grad.data_ptr() >= lo for p in params)    # every .grad aliases a bucket again
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_averager_fixed_order_broadcast_and_reattach_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_robust_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}

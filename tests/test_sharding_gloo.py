"""world_size-2 gloo test of the N>1 host path: shard by graph, one all-gather at the end."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from diffassemble_b200 import sharding


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [4, 6, 3, 5, 2]
    ei, batch = oracle.batch_graphs([oracle.dense_edge_index(n) for n in sizes], sizes)
    x = torch.arange(float(sum(sizes)) * 4).reshape(-1, 4)
    _, _, (x_r,), (n0, n1) = sharding.shard_batch(ei, batch, [x], world, rank)
    counts = []
    for r in range(world):
        g0, g1 = sharding.shard_bounds(len(sizes), world, r)
        counts.append(sum(sizes[g0:g1]))
    poses = x_r * 2.0  # plumbing only: the gather must restore the global node order
    full = sharding.gather_poses(poses, counts)
    ok = torch.equal(full, x * 2.0)
    # the per-rank sampling loop itself (no collective inside it): every rank samples ITS graphs with the CPU oracle
    # (there is no CUDA device here; tests/test_gpu_multirank.py runs the same check through the CUDA path over NCCL)
    # and the gathered trajectory end must equal the unsharded run on the whole batch (V = 0: graphs are independent)
    torch.manual_seed(0)
    ref = oracle.GNNDiffusionRef(steps=20, sampling="DDIM", rotation=True, architecture="exophormer", virt_nodes=0,
                                 model_mean_type=oracle.ModelMeanType.START_X, inference_ratio=5).eval()
    g = torch.Generator().manual_seed(3)
    feats, xT = torch.randn(sum(sizes), 1088, generator=g), torch.randn(sum(sizes), 4, generator=g)
    ei_r, batch_r, (f_r, xs), _ = sharding.shard_batch(ei, batch, [feats, xT], world, rank)
    xu = xT
    with torch.no_grad():
        for i in (15, 10, 5, 0):
            xs, _ = ref.p_sample(xs, torch.full((len(batch_r),), i), i, edge_index=ei_r, patch_feats=f_r, batch=batch_r)
            xu, _ = ref.p_sample(xu, torch.full((len(batch),), i), i, edge_index=ei, patch_feats=feats, batch=batch)
    full = sharding.gather_poses(xs, counts)
    ok = ok and torch.allclose(full, xu, rtol=1e-5, atol=1e-6)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shard_then_single_gather_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffassemble_b200.training import allreduce_gradients

    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    unused = torch.nn.Parameter(torch.zeros(2))  # no grad: must be skipped
    x = torch.full((2, 4), float(rank + 1))
    lin(x).sum().backward()
    allreduce_gradients(list(lin.parameters()) + [unused], world)
    want_w = torch.full((3, 4), 2.0 * (1 + 2) / 2)   # mean over ranks of (2 rows * (rank + 1))
    q.put((rank, torch.allclose(lin.weight.grad, want_w) and torch.allclose(lin.bias.grad, torch.full((3,), 2.0))
           and unused.grad is None))
    dist.destroy_process_group()


def test_gradient_allreduce_world2():
    """N1 multi-GPU leg: DDP-style gradient averaging (one flat all-reduce), checked with gloo on CPU."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]

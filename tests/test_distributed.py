"""world_size-2 gloo test of the multi-GPU plumbing (pair sharding + all-gather of poses/metrics)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import kpreg_b200  # noqa: F401
from kpreg_b200.pipeline import gather_results, shard_pairs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_pairs(n_pairs, rank, world)
    # each "pose row" encodes its pair id so the gathered order can be checked
    local = torch.tensor([[float(i)] * 14 for i in mine], dtype=torch.float32).reshape(len(mine), 14)
    table = gather_results(local, n_pairs, rank, world)
    q.put((rank, table[:, 0].tolist(), table.shape))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharding_and_pose_all_gather_world2():
    world, n_pairs = 2, 7  # uneven: rank 0 has 4 pairs, rank 1 has 3 (padded shard)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ids, shape in results:
        assert tuple(shape) == (n_pairs, 14)
        assert ids == [float(i) for i in range(n_pairs)]  # every rank holds all poses, in pair order

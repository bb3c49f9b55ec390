"""CPU, world_size 2 over gloo: the multi-GPU host logic (scan sharding, view split + peak all-gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvlm_b200.sharding import allgather_peaks, shard_scans, split_views


def test_shard_and_split_arithmetic():
    for world in (1, 2, 3, 8):
        got = sorted(i for r in range(world) for i in shard_scans(256, r, world))
        assert got == list(range(256))
        for v in (8, 100, 200, 7):
            blocks = [split_views(v, r, world) for r in range(world)]
            assert sum(c for _, c in blocks) == v
            assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(5 * n_views * 3, dtype=torch.float32).view(5, n_views, 3)
    s, c = split_views(n_views, rank, world)
    out = allgather_peaks(full[:, s:s + c].contiguous(), n_views)
    q.put((rank, bool(torch.equal(out, full))))
    dist.destroy_process_group()


def test_allgather_peaks_world2_uneven():
    ctx = mp.get_context("spawn")
    for n_views in (8, 7):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert sorted(res) == [(0, True), (1, True)]

"""CPU, world_size 2 over gloo: the multi-GPU host logic (scan sharding, view split + peak all-gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvlm_b200.sharding import allgather_bytes, allgather_peaks, shard_scans, split_views


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


def _rows_worker(rank, world, port, n_views, q):
    """the in-place key gather + row table: what peaks_from_gathered_keys does on the device, checked on CPU"""
    from mvlm_b200.sharding import _view_rows, allgather_keys

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_lm = 5
    slot = split_views(n_views, 0, world)[1]
    full = torch.arange(n_views * n_lm, dtype=torch.int64).view(n_views, n_lm) + 7
    kb = torch.zeros((world, slot, n_lm), dtype=torch.int64)
    s, c = split_views(n_views, rank, world)
    if c:
        kb[rank, :c] = full[s:s + c]
    allgather_keys(kb)
    got = kb.view(world * slot, n_lm).index_select(0, _view_rows(n_views, world, "cpu"))
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


def test_allgather_keys_in_place_world3_and_more_ranks_than_views():
    ctx = mp.get_context("spawn")
    for world, n_views in ((3, 8), (3, 2), (2, 7)):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_rows_worker, args=(r, world, port, n_views, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert sorted(res) == [(r, True) for r in range(world)], (world, n_views, res)


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


def _bytes_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(3)
    ok = True
    # sizes that do not divide by world * 16, a tiny array (some ranks own nothing), uint8 / int32 / float32
    for a in (rng.rand(1001, 3).astype(np.float32), rng.randint(0, 1 << 30, (777, 3)).astype(np.int32),
              rng.randint(0, 256, (33, 17, 3)).astype(np.uint8), np.arange(5, dtype=np.float32)):
        out = allgather_bytes(a, "cpu")
        ok = ok and out.dtype == torch.from_numpy(a).dtype and tuple(out.shape) == a.shape and np.array_equal(out.numpy(), a)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_upload_reassembles_arrays():
    """Each rank contributes 1/world of the bytes; every rank ends with the whole array (view-split mesh upload)."""
    ctx = mp.get_context("spawn")
    for world in (2, 3):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_bytes_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert sorted(res) == [(r, True) for r in range(world)]

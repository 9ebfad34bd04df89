"""CPU, world_size 2 over gloo: stream sharding + the single stats all-gather of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cogstream_b200 import STATS_FIELDS, gather_stats, pad_rows, shard_streams


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_streams(n_streams, rank, world)
        per_rank = (n_streams + world - 1) // world
        rows = []
        for s in mine:  # fake per-stream stats: T=100+s, n_kept=7*s
            r = torch.zeros(len(STATS_FIELDS), dtype=torch.int64)
            r[0], r[5] = 100 + s, 7 * s
            rows.append(r)
        allr = gather_stats(pad_rows(rows, per_rank, "cpu"))
        q.put((rank, allr.clone()))
    finally:
        dist.destroy_process_group()


def _run_world(world, n_streams):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        got = dict(q.get(timeout=100) for _ in range(world))
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    return got


@pytest.mark.timeout(300)
def test_stats_all_gather_world2():
    world, n_streams = 2, 5
    got = None
    for attempt in range(3):  # the rendezvous port is picked by bind(0) + close: another process can grab it in between
        try:
            got = _run_world(world, n_streams)
            break
        except Exception:
            if attempt == 2:
                raise
    assert torch.equal(got[0], got[1])
    allr = got[0]
    assert allr.shape == (2, 3, len(STATS_FIELDS))
    seen = {}
    for r in range(world):
        for i, s in enumerate(shard_streams(n_streams, r, world)):
            seen[s] = allr[r, i]
    assert sorted(seen) == list(range(n_streams))
    for s, row in seen.items():
        assert int(row[0]) == 100 + s and int(row[5]) == 7 * s
    assert int(allr[1, 2].sum()) == 0  # padding row of the rank with fewer streams


def test_gather_stats_without_process_group():
    rows = torch.arange(16).view(2, 8)
    assert torch.equal(gather_stats(rows), rows.unsqueeze(0))

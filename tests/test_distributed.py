"""CPU: the N>1 path (tile sharding + the single all-gather) with world_size-2 gloo processes."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from keep_b200 import distributed as kd


@pytest.mark.parametrize("n,ws", [(10, 1), (10, 2), (10, 4), (7, 8), (50000, 8), (0, 2), (3, 8)])
def test_shard_ranges_partition_the_tiles(n, ws):
    ranges = [kd.shard_range(n, r, ws) for r in range(ws)]
    covered = [i for lo, hi in ranges for i in range(lo, hi)] if n < 1000 else None
    if covered is not None:
        assert covered == list(range(n))
    assert sum(hi - lo for lo, hi in ranges) == n
    assert all(hi - lo <= kd.shard_size(n, ws) for lo, hi in ranges)
    assert all(ranges[i][1] == ranges[i + 1][0] or ranges[i + 1][0] == n for i in range(ws - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _FakeModel:
    """Stands in for KEEPModel in the gloo test: a deterministic per-tile 'embedding' (no CUDA needed)."""

    class config:
        projection_dim = 8

    logit_scale = torch.zeros(())

    def encode_image(self, tiles):
        return tiles.reshape(tiles.shape[0], -1)[:, :8] * 2.0


def _worker(rank, ws, port, n_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    r, _, w = kd.init_from_env(backend="gloo")
    assert (r, w) == (rank, ws)
    all_tiles = torch.arange(n_total * 16, dtype=torch.float32).reshape(n_total, 1, 4, 4)
    seen = []

    def load(lo, hi):
        seen.append((lo, hi))
        return all_tiles[lo:hi]

    full = kd.encode_tiles_sharded(_FakeModel(), n_total, load, batch=3)
    expect = all_tiles.reshape(n_total, -1)[:, :8] * 2.0
    ok = torch.equal(full, expect)
    lo, hi = kd.shard_range(n_total, rank, ws)
    ok = ok and (not seen or (seen[0][0] == lo and seen[-1][1] == hi))
    half = kd.encode_tiles_sharded(_FakeModel(), n_total, load, batch=4, gather_dtype=torch.float16)
    ok = ok and torch.equal(half, expect.half().float())
    t = kd.barrier_max_ms(float(rank + 1), torch.device("cpu"))
    ok = ok and t == float(ws)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [11, 2, 1])
def test_sharded_encode_and_allgather_gloo_world2(n_total):
    ws = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(ws, port, n_total, ret), nprocs=ws, join=True)
    assert dict(ret) == {0: True, 1: True}

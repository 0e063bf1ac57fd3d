"""CPU, world_size 2 over gloo: the stream-sharding host logic used for N > 1 GPUs (no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepfilternet_b200.sharding import enhance_sharded, gather_rank_rows, shard_range, shard_sizes


def test_shard_range_partitions_exactly():
    for n in (0, 1, 2, 7, 128, 129, 4096):
        for w in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
            sizes = shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_gather_rank_rows_without_process_group():
    assert gather_rank_rows([1, 2.5]) == [[1.0, 2.5]]


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
    audio = torch.arange(n_streams * 6, dtype=torch.float32).reshape(n_streams, 6)
    calls = []

    def fake_enhance(x):  # stands in for enhance(): per-stream independent, records what it was given
        calls.append(tuple(x.shape))
        return x * 2.0 + 1.0

    out = enhance_sharded(fake_enhance, audio, gather_to=0)
    local = enhance_sharded(fake_enhance, audio)
    s, e = shard_range(n_streams, rank, world)
    ok = torch.equal(local, audio[s:e] * 2.0 + 1.0) and calls[0][0] == e - s
    if rank == 0:
        ok = ok and out is not None and torch.equal(out, audio * 2.0 + 1.0)
    else:
        ok = ok and out is None
    # per-GPU report of bench.py: every rank gets every rank's row
    rows = gather_rank_rows([rank, 10.0 + rank, float("nan")])
    ok = ok and len(rows) == world and all(rows[r][0] == r and rows[r][1] == 10.0 + r and rows[r][2] != rows[r][2] for r in range(world))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [5, 8])
def test_two_rank_sharded_enhance_over_gloo(n_streams):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]

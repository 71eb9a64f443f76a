"""N > 1 host logic on CPU: world_size-2 `gloo` process group (no GPU, no kernels).

Shards a batch of utterances over two ranks, lets each rank produce the durations of ITS utterances with
the CPU oracle (stand-in for the device op -- the thing under test is the sharding / gather plumbing of
face_gan_tts_b200.sharding), all-gathers them and checks every rank ends up with the durations of the
whole batch in the original order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from face_gan_tts_b200 import sharding, synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _durations(value, t_x, t_y):
    path = np.zeros(value.shape, np.int32)
    oracle.maximum_path_c(path, np.ascontiguousarray(value, np.float32).copy(), t_x.astype(np.int32), t_y.astype(np.int32))
    return torch.from_numpy(path.sum(-1).astype(np.int32))


def _worker(rank, world, port, B, uneven, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        v, t_x, t_y = synthetic.mas_value(B, 24, 60, seed=7, tx_lo=5, ty_lo=30)
        v, t_x, t_y = v.numpy(), t_x.numpy(), t_y.numpy()
        full = _durations(v, t_x, t_y)
        n = B if not uneven else B - 1
        lo, hi = sharding.shard_range(n, world, rank)
        local = _durations(v[lo:hi], t_x[lo:hi], t_y[lo:hi])
        got = sharding.all_gather_durations(local)
        assert got.shape == (n, 24) and torch.equal(got, full[:n]), "gathered durations differ from the whole batch"
        out, work = sharding.all_gather_durations(local, async_op=True)
        if work is not None:
            work.wait()
        assert torch.equal(out, full[:n])
        q.put((rank, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("uneven", [False, True])
def test_shard_and_gather_durations_world2(uneven):
    world, B = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, uneven, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    spans = sorted(q.get(timeout=5)[1:] for _ in range(world))
    n = B - 1 if uneven else B
    assert spans[0][0] == 0 and spans[-1][1] == n and spans[0][1] == spans[1][0]


def test_shard_range_partitions():
    for n in (0, 1, 7, 32, 1024):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_shards_cover_and_balance():
    g = torch.Generator().manual_seed(3)
    t_x = torch.randint(20, 190, (64,), generator=g).tolist()
    t_y = [max(a, b) for a, b in zip(t_x, torch.randint(200, 1000, (64,), generator=g).tolist())]
    for world in (2, 4, 8):
        shards = sharding.balanced_shards(t_x, t_y, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(64))
        loads = [sum(t_x[i] * t_y[i] for i in s) for s in shards]
        assert max(loads) <= 1.15 * (sum(loads) / world)

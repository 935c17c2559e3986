"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from warprnnt_pytorch.distributed import reduce_loss, shard_bounds
    n_global = 13
    start, cnt = shard_bounds(n_global, rank, world)
    costs = torch.arange(n_global, dtype=torch.float32) * 1.5 + 2.0       # per-utterance costs
    local = costs[start:start + cnt].sum()
    loss_mean, n = reduce_loss(local, cnt, 'mean')
    loss_sum, _ = reduce_loss(local, cnt, 'sum')
    q.put((rank, start, cnt, float(loss_mean), float(loss_sum), n))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    import sys
    from warprnnt_pytorch.distributed import shard_bounds
    for n in (1, 7, 128, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert shard_bounds(1024, 3, 8) == (384, 128)        # BASELINE config 5: 128 utterances per GPU
    with pytest.raises(ValueError):
        shard_bounds(8, 8, 8)


def test_reduce_loss_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    costs = torch.arange(13, dtype=torch.float32) * 1.5 + 2.0
    assert [o[1:3] for o in out] == [(0, 7), (7, 6)]
    for o in out:
        assert abs(o[3] - float(costs.mean())) < 1e-5      # identical global mean on every rank
        assert abs(o[4] - float(costs.sum())) < 1e-4
        assert o[5] == 13


def test_reduce_loss_single_process():
    from warprnnt_pytorch.distributed import reduce_loss
    loss, n = reduce_loss(torch.tensor(12.0), 4, 'mean')
    assert float(loss) == 3.0 and n == 4
    with pytest.raises(ValueError):
        reduce_loss(torch.tensor(1.0), 1, 'none')

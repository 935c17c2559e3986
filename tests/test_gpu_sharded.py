"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped on a 1-GPU box): two NCCL ranks each
own half the batch; the global loss and the per-shard gradients must equal the single-GPU result."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, acts_np, labels_np, tl_np, ul_np, reduction, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from warprnnt_pytorch.distributed import ShardedRNNTLoss, shard_bounds
    N = acts_np.shape[0]
    s, c = shard_bounds(N, rank, world)
    acts = torch.tensor(acts_np[s:s + c], device=dev, requires_grad=True)
    # every shard keeps the global tensor extents (maxT, maxU), as a sharded joint network would
    loss = ShardedRNNTLoss(reduction=reduction, n_global=N)
    from warprnnt_pytorch import certify_inputs  # noqa: F401
    labels, tl, ul = (torch.as_tensor(x[s:s + c]).to(dev) for x in (labels_np, tl_np, ul_np))
    # certify_inputs wants max(len) == extent on every shard: give each shard one full-length utterance
    out = loss(acts, labels, tl, ul)
    out.backward()
    q.put((rank, s, c, float(out.item()), acts.grad.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_two_rank_shard_equals_single_gpu(reduction):
    import torch.multiprocessing as mp
    from oracle import pyoracle
    rng = np.random.default_rng(0)
    N, T, U, V = 6, 14, 5, 28
    acts = rng.standard_normal((N, T, U, V)).astype(np.float32)
    labels = rng.integers(1, V, size=(N, U - 1)).astype(np.int32)
    tl = rng.integers(T // 2, T + 1, size=N).astype(np.int32)
    ul = rng.integers(0, U, size=N).astype(np.int32)
    tl[0] = tl[3] = T
    ul[0] = ul[3] = U - 1
    c_ref, g_ref, _ = pyoracle.rnnt_logits(acts.astype(np.float64), labels, tl, ul, 0)
    want = c_ref.sum() / (N if reduction == "mean" else 1)
    g_want = g_ref / (N if reduction == "mean" else 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, acts, labels, tl, ul, reduction, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted((q.get(timeout=300) for _ in range(2)), key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, s, c, loss, g in outs:
        assert abs(loss - want) < 1e-4 * abs(want), (rank, loss, want)
        assert np.allclose(g, g_want[s:s + c], rtol=1e-4, atol=1e-6)

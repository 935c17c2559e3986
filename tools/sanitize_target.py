"""Small mixed-shape workload for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "warp-transducer_b200"))
import warprnnt_pytorch.warp_rnnt as wr  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
for (N, T, U, V) in [(3, 9, 5, 28), (2, 12, 40, 6), (2, 7, 4, 301), (2, 5, 3, 1028), (1, 3, 2, 8200), (9, 30, 70, 50)]:
    acts = torch.tensor(rng.standard_normal((N, T, U, V)).astype(np.float32), device=dev)
    labels = torch.as_tensor(rng.integers(1, V, size=(N, U - 1)).astype(np.int32)).to(dev)
    tl = torch.as_tensor(rng.integers(T // 2 + 1, T + 1, size=N).astype(np.int32)).to(dev)
    ul = torch.as_tensor(rng.integers(0, U, size=N).astype(np.int32)).to(dev)
    costs = torch.empty(N, device=dev)
    grads = torch.empty_like(acts)
    os.environ.pop("RNNT_B200_GROUPS", None)
    ws = wr.gpu_rnnt_async(acts, labels, tl, ul, costs, grads, 0)
    ws2 = wr.gpu_rnnt_forward(acts, labels, tl, ul, costs, 0, True)
    wr.gpu_rnnt_backward(acts, labels, tl, ul, grads, None, 0, 0.5, ws2)
    torch.cuda.synchronize()
    print((N, T, U, V), "cost0", float(costs[0]), "finite", bool(torch.isfinite(grads).all()))

# 16-bit storage and the additive-joint variant
from warprnnt_pytorch import RNNTLoss  # noqa: E402
from warprnnt_pytorch.joint import AddJointRNNTLoss  # noqa: E402
for dt in (torch.bfloat16, torch.float16):
    N, T, U, V = 2, 9, 5, 512
    acts = torch.tensor(rng.standard_normal((N, T, U, V)).astype(np.float32), device=dev).to(dt).requires_grad_(True)
    labels = torch.as_tensor(rng.integers(1, V, size=(N, U - 1)).astype(np.int32)).to(dev)
    tl = torch.tensor([T, 6], dtype=torch.int32, device=dev)
    ul = torch.tensor([U - 1, 2], dtype=torch.int32, device=dev)
    RNNTLoss()(acts, labels, tl, ul).backward()
    torch.cuda.synchronize()
    print(dt, "finite", bool(torch.isfinite(acts.grad.float()).all()))
for (N, T, U, V) in [(2, 9, 5, 28), (2, 20, 34, 700), (2, 40, 7, 130), (2, 70, 21, 520)]:
    trans = torch.tensor(rng.standard_normal((N, T, V)).astype(np.float32), device=dev, requires_grad=True)
    pred = torch.tensor(rng.standard_normal((N, U, V)).astype(np.float32), device=dev, requires_grad=True)
    labels = torch.as_tensor(rng.integers(1, V, size=(N, U - 1)).astype(np.int32)).to(dev)
    tl = torch.tensor([T, T - 3], dtype=torch.int32, device=dev)
    ul = torch.tensor([U - 1, 1], dtype=torch.int32, device=dev)
    AddJointRNNTLoss()(trans, pred, labels, tl, ul).backward()
    torch.cuda.synchronize()
    print("add-joint", (N, T, U, V), "finite", bool(torch.isfinite(trans.grad).all() and torch.isfinite(pred.grad).all()))

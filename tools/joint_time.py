"""Additive joint: AddJointRNNTLoss on (trans, pred) vs materialising acts = trans + pred and running
the dense RNNTLoss (what pytorch_binding/test/test_time.py:73 does), forward + backward."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "warp-transducer_b200"))
from warprnnt_pytorch import RNNTLoss  # noqa: E402
from warprnnt_pytorch.joint import AddJointRNNTLoss  # noqa: E402

dev = torch.device("cuda:0")
for (N, T, L, V) in [(128, 150, 20, 5000), (128, 150, 40, 28), (64, 1500, 300, 50)]:
    U = L + 1
    trans = torch.rand((N, T, V), device=dev, requires_grad=True)
    pred = torch.rand((N, U, V), device=dev, requires_grad=True)
    labels = torch.as_tensor(np.random.default_rng(1).integers(1, V, size=(N, L)).astype(np.int32)).to(dev)
    tl = torch.full((N,), T, dtype=torch.int32, device=dev)
    ul = torch.full((N,), L, dtype=torch.int32, device=dev)
    fused, dense = AddJointRNNTLoss(), RNNTLoss()

    def run_fused():
        trans.grad = pred.grad = None
        fused(trans, pred, labels, tl, ul).backward()

    def run_dense():
        trans.grad = pred.grad = None
        acts = trans.unsqueeze(2) + pred.unsqueeze(1)
        dense(acts, labels, tl, ul).backward()

    res = {}
    for name, fn in (("fused add-joint", run_fused), ("materialise + dense", run_dense)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10
        g = (trans.grad.clone(), pred.grad.clone())
        res[name + " grads"] = g
    d1 = (res["fused add-joint grads"][0] - res["materialise + dense grads"][0]).abs().max().item()
    print("N=%d T=%d U=%d V=%d: fused %.3f ms/step (%.0f utt/s)  vs materialise+dense %.3f ms/step  (x%.1f)  max|dgrad| %.2e"
          % (N, T, U, V, res["fused add-joint"], N / res["fused add-joint"] * 1e3, res["materialise + dense"],
             res["materialise + dense"] / res["fused add-joint"], d1), flush=True)

"""Target for ncu captures: N iterations of the async C-ABI on one BASELINE shape, nothing else.
    ncu --set full -k regex:'rowstats|lattice|grad_kernel' -s 3 -c 3 python tools/profile_target.py c3 2
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "warp-transducer_b200"))
import warprnnt_pytorch.warp_rnnt as wr  # noqa: E402

CFG = {"c2": (128, 150, 40, 28), "c3": (128, 150, 20, 5000), "c4": (64, 1500, 300, 50),
       "c5": (128, 200, 40, 5000), "c3s": (16, 150, 20, 5000)}
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dt = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[3] if len(sys.argv) > 3 else "fp32"]
N, T, L, V = CFG[name]
dev = torch.device("cuda:0")
acts = torch.rand((N, T, L + 1, V), device=dev).to(dt)
grads = torch.empty_like(acts)
labels = torch.as_tensor(np.random.default_rng(1).integers(1, V, size=(N, L)).astype(np.int32)).to(dev)
tl = torch.full((N,), T, dtype=torch.int32, device=dev)
ul = torch.full((N,), L, dtype=torch.int32, device=dev)
costs = torch.empty(N, device=dev)
ws = None
for _ in range(iters):
    ws = wr.gpu_rnnt_async(acts, labels, tl, ul, costs, grads, 0, 1.0, ws)
torch.cuda.synchronize()
print(name, "cost0", costs[0].item())

import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "warp-transducer_b200"))
import warprnnt_pytorch.warp_rnnt as wr
from oracle import pyoracle
N, T, U, V = [int(a) for a in sys.argv[1:5]]
rng = np.random.default_rng(0)
acts = rng.standard_normal((N, T, U, V)).astype(np.float32)
labels = rng.integers(1, V, size=(N, max(U - 1, 1))).astype(np.int32)
tl = np.full(N, T, np.int32); ul = np.full(N, U - 1, np.int32)
dev = torch.device("cuda:0")
a = torch.tensor(acts, device=dev); g = torch.full_like(a, float("nan")); c = torch.empty(N, device=dev)
wr.gpu_rnnt_async(a, torch.tensor(labels, device=dev), torch.tensor(tl, device=dev), torch.tensor(ul, device=dev), c, g, 0)
torch.cuda.synchronize()
cr, gr, _ = pyoracle.rnnt_logits(acts.astype(np.float64), labels, tl, ul, 0)
g = g.cpu().numpy()
print("costs", c.cpu().numpy(), cr)
print("g[0,0,0,:8]", g[0, 0, 0, :8], "ref", gr[0, 0, 0, :8])
print("nonfinite", np.count_nonzero(~np.isfinite(g)), "of", g.size, "maxdiff", np.nanmax(np.abs(g - gr)))
bad = ~np.isfinite(g)
for r in range(min(3, T * U)):
    t, u = divmod(r, U)
    idx = np.nonzero(bad[0, t, u])[0]
    print("row", r, "nonfinite count", idx.size, "first", idx[:6], "last", idx[-6:] if idx.size else [])

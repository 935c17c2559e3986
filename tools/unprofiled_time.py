import sys, os, numpy as np, torch
sys.path.insert(0, "warp-transducer_b200")
import warprnnt_pytorch.warp_rnnt as wr
# un-profiled timing (PDL active) of c2 / c4 with different RPT / PDL settings is done per process via env
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
want = sys.argv[1:]
for name, (N, T, L, V) in {"c2": (128, 150, 40, 28), "c4": (64, 1500, 300, 50), "c3": (128, 150, 20, 5000)}.items():
    if want and name not in want:
        continue
    U = L + 1
    acts = torch.rand((N, T, U, V), device=dev); grads = torch.empty_like(acts)
    labels = torch.as_tensor(np.random.default_rng(1).integers(1, V, size=(N, L)).astype(np.int32)).to(dev)
    tl = torch.full((N,), T, dtype=torch.int32, device=dev); ul = torch.full((N,), L, dtype=torch.int32, device=dev)
    costs = torch.empty(N, device=dev); ws = None
    ts = []
    for it in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ws = wr.gpu_rnnt_async(acts, labels, tl, ul, costs, grads, 0, 1.0, ws); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("unprofiled %s: %.4f ms (min %.4f) (PDL=%s NT=%s GROUPS=%s)" % (name, float(np.median(ts[4:])), float(np.min(ts[4:])), os.environ.get("RNNT_B200_PDL", "0"), os.environ.get("RNNT_B200_CHUNK_NT", "256"), os.environ.get("RNNT_B200_GROUPS", "auto")), flush=True)
    del acts, grads

"""Quick device-side timing of the async C-ABI on the BASELINE shapes (dev tool, not the bench)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "warp-transducer_b200"))
import warprnnt_pytorch.warp_rnnt as wr  # noqa: E402

CFG = {"c2": (128, 150, 40, 28), "c3": (128, 150, 20, 5000), "c4": (64, 1500, 300, 50),
       "c5": (128, 200, 40, 5000), "c3h": (64, 150, 20, 5000), "c3odd": (128, 150, 20, 5001), "c3v2": (128, 150, 20, 5002),
       "v1025": (64, 150, 40, 1025), "v4097": (64, 150, 40, 4097)}


def main():
    dt = torch.float32
    if "--fp64" in sys.argv:
        sys.argv.remove("--fp64")
        dt = torch.float64
    if "--bf16" in sys.argv:
        sys.argv.remove("--bf16")
        dt = torch.bfloat16
    names = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in names:
        N, T, L, V = CFG[name]
        U = L + 1
        acts = torch.rand((N, T, U, V), device=dev).to(dt)
        grads = torch.empty_like(acts)
        rng = np.random.default_rng(1)
        labels = torch.as_tensor(rng.integers(1, V, size=(N, L)).astype(np.int32)).to(dev)
        tl = torch.full((N,), T, dtype=torch.int32, device=dev)
        ul = torch.full((N,), L, dtype=torch.int32, device=dev)
        costs = torch.empty(N, device=dev, dtype=torch.float64 if dt == torch.float64 else torch.float32)
        ws = None
        for mode, g in (("loss+grad", grads), ("loss", None)):
            ts = []
            wr.set_profiling(True)
            km = np.zeros(3)
            for it in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ws = wr.gpu_rnnt_async(acts, labels, tl, ul, costs, g, 0, 1.0, ws)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
                if it >= 3:
                    km += np.array(wr.last_kernel_ms()) / 5
            t = float(np.median(ts[3:]))
            E = N * T * U * V
            bytes_ = (12 if g is not None else 4) * E * acts.element_size() // 4
            print("%s %-9s N=%d T=%d U=%d V=%d: %.3f ms  %.0f utt/s  %.0f GB/s (algorithmic %d B/elt)  cost[0]=%.3f  kernels(rowstats,lattice,grad)=%s" % (
                name, mode, N, T, U, V, t, N / t * 1e3, bytes_ / t / 1e6, 12 if g is not None else 4,
                costs[0].item(), np.round(km, 3)), flush=True)
        del acts, grads
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

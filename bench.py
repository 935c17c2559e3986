#!/usr/bin/env python
"""bench.py — RNN-T loss+grad throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus 1 --steps 20 --warmup 3                      # our arm, 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W       # our arm, N GPUs (weak scaling)
    python bench.py --impl reference --gpus 1 --steps K --warmup W      # reference CPU path on host cores

A step = one pass of the hot path (log-softmax statistics -> alpha/beta lattice -> dense gradient)
over one batch of synthetic logits; workload = BASELINE config "N=128, T=150, L=20, A=5000 fp32"
per GPU (the configuration the metric is quoted on).  Multi-GPU: every rank owns its own
128-utterance shard (utterances are independent), one NCCL all-reduce of the scalar loss per step.

value      device-resident: inputs already in HBM, compute_rnnt_loss_async + loss all-reduce,
           K steps between CUDA events, max over ranks.
e2e        through the reference-facing C-ABI call compute_rnnt_loss() with the step's inputs
           copied from pinned host memory inside the timed region and the costs landing on the host.
roofline   the dominant kernel (grad_row_kernel, 8 B/element algorithmic) timed with CUDA events on the
           library's own stream during the timed steps, against MEASURED_PEAKS.json.
cpu_baseline  the reference's CPU path (oracle/_ref, else the oracle port) on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "warp-transducer_b200"))

WORKLOADS = {   # name: (N per GPU, T, L, V)   BASELINE.json configs
    "c2": (128, 150, 40, 28),
    "c3": (128, 150, 20, 5000),
    "c4": (64, 1500, 300, 50),
    "c5": (128, 200, 40, 5000),     # config 5 = 1024 utterances / 8 GPUs
}
METRIC = "RNN-T loss+grad utterances/s at T=150,L=20,A=5000"
UNIT = "utterances/s"


def gen_labels(V, L, N):
    """Labels in [1, V-1] with forced repeats, the same row for every utterance — the recipe of
    the reference harness (tests/random.cpp:22-38, tests/test_time.cu:40-43), numpy generator."""
    rng = np.random.default_rng(1)
    lab = rng.integers(1, V, size=L).astype(np.int32)
    if L >= 3:
        lab[L // 2] = lab[L // 2 + 1]
        lab[L // 2 - 1] = lab[L // 2]
    return np.tile(lab, (N, 1))


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.maxc, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxc = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.maxc, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.maxc,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU implementation of the path on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference_step_fn(V, T, L, n_utt, threads):
    """Returns (fn, kind): fn() runs logits -> loss + dense logits-gradient for n_utt utterances of
    the workload shape on the CPU.  kind 'reference': torch.log_softmax -> oracle/_ref
    compute_rnnt_loss(loc=CPU, OpenMP) -> log-softmax backward, i.e. exactly what the reference's
    warprnnt_pytorch composes on CPU (__init__.py:95-98 + autograd).  kind 'port': the oracle."""
    import torch
    from oracle import pyoracle
    U = L + 1
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(0)
    acts = torch.rand((n_utt, T, U, V), generator=gen, dtype=torch.float32)
    labels = gen_labels(V, L, n_utt)
    tl = np.full(n_utt, T, np.int32)
    ul = np.full(n_utt, L, np.int32)
    if pyoracle.have_ref_cpu():
        import ctypes as C
        lib = pyoracle.load_ref_cpu()
        fn = lib.compute_rnnt_loss
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p, pyoracle.RnntOptions]
        nbytes = C.c_size_t(0)
        lib.get_workspace_size.argtypes = [C.c_int, C.c_int, C.c_int, C.c_bool, C.POINTER(C.c_size_t), C.c_size_t]
        lib.get_workspace_size(T, U, n_utt, False, C.byref(nbytes), 4)
        ws = np.zeros(nbytes.value, np.uint8)
        costs = np.zeros(n_utt, np.float32)
        g = torch.empty_like(acts)
        opt = pyoracle.RnntOptions(loc=0, num_threads=threads, stream=None, blank_label=0, maxT=T,
                                   maxU=U, batch_first=True)

        def step():
            lp = torch.log_softmax(acts, -1)
            rc = fn(lp.data_ptr(), g.data_ptr(), labels.ctypes.data, ul.ctypes.data, tl.ctypes.data,
                    V, n_utt, costs.ctypes.data, ws.ctypes.data, opt)
            assert rc == 0
            dx = g - torch.exp(lp) * g.sum(-1, keepdim=True)
            return float(costs.sum()), dx
        return step, "reference"

    acts_np = acts.numpy()

    def step():
        c, dx, _ = pyoracle.rnnt_logits(acts_np, labels, tl, ul, 0, True, threads)
        return float(c.sum()), dx
    return step, "port"


def time_cpu(V, T, L, threads, target_s, max_utt):
    """Pick a bounded sample (about target_s seconds of CPU work) and time it."""
    pilot_n = min(4, max_utt)
    step, kind = cpu_reference_step_fn(V, T, L, pilot_n, threads)
    step()
    t0 = time.perf_counter()
    step()
    per_utt = (time.perf_counter() - t0) / pilot_n
    n = int(max(pilot_n, min(max_utt, target_s / max(per_utt, 1e-9))))
    # the reference parallelises over utterances (cpu_rnnt.h:290 `#pragma omp parallel for`), so a tiny
    # sample would leave cores idle and understate the CPU; measured on the 128-core GPU box the rate
    # is flat from 40 utterances up (11.4 @40, 11.9 @128 utt/s), so 32 is the floor
    n = min(max(n, min(32, threads)), max_utt)
    step, kind = cpu_reference_step_fn(V, T, L, n, threads)
    step()
    return step, kind, n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, T, L, V = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    # size the per-step sample so that the whole --steps/--warmup run stays within ~2.5 minutes
    step, kind, n = time_cpu(V, T, L, cores, target_s=150.0 / (args.steps + args.warmup), max_utt=N)
    for _ in range(max(args.warmup - 1, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt
    sample = "%d of %d utterances of the workload per step, shape T=%d U=%d V=%d" % (n, N, T, L + 1, V)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.workload, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def needs_no_flush(name):
    N, T, L, V = WORKLOADS[name]
    return N * T * (L + 1) * V * 4 > (1 << 30)


def workload_config(name, world):
    N, T, L, V = WORKLOADS[name]
    return {"workload": "%s: N=%d per GPU, T=%d, L=%d (U=%d), A=%d, fp32 logits ~U[0,1), full lengths, blank 0"
                        % (name, N, T, L, L + 1, V),
            "global_batch": N * world, "parallelism": "batch-sharded x%d" % world,
            "l2": ("no flush: per-step inputs (%.2f GB logits) exceed the 126 MB L2" % (N * T * (L + 1) * V * 4 / 1e9))
                  if needs_no_flush(name) else "L2 flushed (256 MB write) before every timed step; steps timed individually"}


# ----------------------------------------------------------------------------------------------
# Our arm
# ----------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import warprnnt_pytorch.warp_rnnt as wr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    N, T, L, V = WORKLOADS[args.workload]
    U = L + 1
    E = N * T * U * V
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = torch.rand((N, T, U, V), generator=gen, device=dev, dtype=torch.float32)
    grads = torch.empty_like(acts)
    labels_np = gen_labels(V, L, N)
    labels = torch.as_tensor(labels_np).to(dev)
    tl = torch.full((N,), T, dtype=torch.int32, device=dev)
    ul = torch.full((N,), L, dtype=torch.int32, device=dev)
    costs = torch.empty(N, device=dev)
    ws = torch.empty(wr.workspace_size(T, U, N, 4), dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The scalar-loss all-reduce runs on NCCL's own stream and is consumed one step later, so the
    # compute stream of a rank never stalls on a slower peer inside a step (gradients are local;
    # only the logged loss crosses ranks).  Two alternating buffers keep step k's reduction intact
    # while step k+1 is enqueued.
    loss2 = [torch.zeros(1, device=dev), torch.zeros(1, device=dev)]
    state = {"k": 0, "pending": None}

    def step():
        wr.gpu_rnnt_async(acts, labels, tl, ul, costs, grads, 0, 1.0, ws)
        buf = loss2[state["k"] & 1]
        torch.sum(costs, 0, keepdim=True, out=buf)
        if world > 1 and not os.environ.get("BENCH_NO_ALLREDUCE"):
            if state["pending"] is not None:
                state["pending"].wait()          # step k-1's collective: finished during this step's kernels
            state["pending"] = dist.all_reduce(buf, async_op=True)   # one scalar over NVLink
        state["k"] += 1

    def drain():
        if state["pending"] is not None:
            state["pending"].wait()
            state["pending"] = None

    wr.set_profiling(True)
    # everything with variable host cost (NVML init, thread start, event creation) happens BEFORE the
    # barrier: ranks must leave it aligned, a late starter is waited for by all the others through
    # the loss all-reduce and with K ~ 20 steps of 3.5 ms a 20 ms skew is a 30 % error
    sampler = ClockSampler(local)
    sampler.start()
    kms = np.zeros(3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    wr.profile_collect()                              # drop the warm-up records
    barrier()
    if needs_no_flush(args.workload):
        e0.record()
        for _ in range(args.steps):
            step()                                    # no host synchronisation inside the timed region
        drain()                                       # the last step's all-reduce is inside the timed region
        e1.record()
        barrier()
        total_ms = e0.elapsed_time(e1)
        kms = np.array(wr.profile_collect()[1]) * args.steps
    else:   # working set fits the L2: flush it before every step and time the steps one by one
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        total_ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            e0.record()
            step()
            drain()
            e1.record()
            kms += np.array(wr.last_kernel_ms())
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        barrier()
    launches = wr.last_launch_count() * args.steps
    wr.set_profiling(False)

    # ---- secondary, reported separately: ragged lengths ~U[0.5,1]*max, seed 2.
    # Padded cells are not read (pass 1 skips them, pass 2 writes zeros), so bytes move less.
    ragged = None
    try:
        rng = np.random.default_rng(2)
        tl_r = torch.as_tensor(np.maximum(1, (rng.uniform(0.5, 1.0, N) * T)).astype(np.int32)).to(dev)
        ul_r = torch.as_tensor((rng.uniform(0.5, 1.0, N) * L).astype(np.int32)).to(dev)
        for _ in range(3):
            wr.gpu_rnnt_async(acts, labels, tl_r, ul_r, costs, grads, 0, 1.0, ws)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        r0.record()
        rsteps = max(3, args.steps // 2)
        for _ in range(rsteps):
            wr.gpu_rnnt_async(acts, labels, tl_r, ul_r, costs, grads, 0, 1.0, ws)
        r1.record()
        torch.cuda.synchronize()
        rms = r0.elapsed_time(r1) / rsteps
        valid = float((tl_r.double() * (ul_r.double() + 1)).sum().item())
        ragged = {"ms_per_step": rms, "value": N / (rms * 1e-3), "unit": UNIT + " per GPU",
                  "lengths": "T_b, L_b ~ U[0.5,1] x max (seed 2)", "valid_cell_fraction": valid / (N * T * U),
                  "algorithmic_GBps": (8.0 * valid * V + 4.0 * N * T * U * V) / (rms * 1e-3) / 1e9}
    except Exception as ex:
        ragged = {"error": repr(ex)[:200]}

    # ---- end to end through compute_rnnt_loss(): pinned host inputs -> device, costs -> host
    acts_host = torch.empty((N, T, U, V), dtype=torch.float32, pin_memory=True)
    acts_host.copy_(acts)
    labels_host = torch.as_tensor(labels_np).pin_memory()
    tl_host = torch.full((N,), T, dtype=torch.int32).pin_memory()
    ul_host = torch.full((N,), L, dtype=torch.int32).pin_memory()
    costs_host = torch.zeros(N, dtype=torch.float32).pin_memory()
    opt = wr.rnntOptions(loc=1, num_threads=0, stream=torch.cuda.current_stream().cuda_stream,
                         blank_label=0, maxT=T, maxU=U, batch_first=True)

    def e2e_step():
        acts.copy_(acts_host, non_blocking=True)
        labels.copy_(labels_host, non_blocking=True)
        tl.copy_(tl_host, non_blocking=True)
        ul.copy_(ul_host, non_blocking=True)
        st = wr.lib().compute_rnnt_loss(acts.data_ptr(), grads.data_ptr(), labels.data_ptr(),
                                        ul.data_ptr(), tl.data_ptr(), V, N, costs_host.data_ptr(),
                                        ws.data_ptr(), opt)      # returns with costs on the host
        assert st == 0
        tot = float(costs_host.sum())
        if world > 1:
            t = torch.tensor([tot], device=dev)
            dist.all_reduce(t)
            tot = float(t.item())
        return tot

    e2e_steps = max(3, min(args.steps, 10))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_step()
    barrier()
    f0.record()
    for _ in range(e2e_steps):
        e2e_step()
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    clocks = sampler.result()

    per_rank = None
    if world > 1:
        # diagnostics: every rank's own kernel times and loop time (gathered, not used for `value`)
        mine = torch.tensor([total_ms, kms[0] / args.steps, kms[1] / args.steps, kms[2] / args.steps],
                            device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(v), 4) for v in r.tolist()] for r in allr]
        t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = t.tolist()
    ms_per_step = total_ms / args.steps
    value = N * world / (ms_per_step * 1e-3)
    e2e_value = N * world / (e2e_ms / e2e_steps * 1e-3)
    h2d = acts_host.numel() * 4 + labels_host.numel() * 4 + 2 * N * 4
    d2h = N * 4

    peaks, peak_src = measured_peaks()
    kms /= args.steps
    grad_ms, rows_ms, lat_ms = float(kms[2]), float(kms[0]), float(kms[1])
    achieved = 8.0 * E / (grad_ms * 1e-3) / 1e9 if grad_ms > 0 else None
    roofline = {
        "bound": "hbm", "kernel": "grad_row_kernel (pass 2: read logits 4 B + write gradient 4 B per element)",
        "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": achieved / peaks["hbm_gbs"] if achieved else None, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
        "traffic": None, "ms_per_launch": grad_ms,
        "frac_of_8TBps": achieved / 8000.0 if achieved else None,
        "other_kernels": {
            "rowstats_row_kernel": {"ms": rows_ms, "algorithmic_GBps": 4.0 * E / (rows_ms * 1e-3) / 1e9 if rows_ms > 0 else None},
            "lattice_kernel": {"ms": lat_ms, "bound": "latency"},
            "path_12B_per_elt_GBps": 12.0 * E / ((rows_ms + lat_ms + grad_ms) * 1e-3) / 1e9 if grad_ms > 0 else None,
        },
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("grad_kernel_c3_bytes")
        except Exception:
            pass

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "api": "compute_rnnt_loss (C-ABI, host-synchronous); inputs from pinned host memory each step"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "lib": os.path.relpath(wr.lib_path(), ROOT),
        }
        line["ragged_lengths"] = ragged
        if per_rank is not None:
            line["per_rank_ms"] = {"columns": ["loop_total", "rowstats", "lattice", "grad"], "rows": per_rank}
    if world > 1:
        dist.barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        try:
            stepf, kind, n = time_cpu(V, T, L, cores, target_s=6.0, max_utt=N)
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                stepf()
            dt = (time.perf_counter() - t0) / reps
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "%d of %d utterances of the workload (T=%d U=%d V=%d), %d timed passes"
                                              % (n, N, T, U, V, reps)}
        except Exception as ex:   # the baseline is reporting only; never lose the GPU line over it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "unavailable",
                                    "sample": repr(ex)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: relaunch under torchrun when invoked directly with --gpus N
            import subprocess
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                   "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__),
                   "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
                   "--workload", args.workload] + (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
            raise SystemExit(subprocess.call(cmd))
        run_b200_arm(args)


if __name__ == "__main__":
    main()

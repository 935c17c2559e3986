"""warprnnt_pytorch operator surface on the GPU — reads like the reference's own
pytorch_binding/test/test.py (small_test / big_test) plus reduction and autograd semantics."""
import numpy as np
import pytest
import torch

from oracle import pyoracle

pytestmark = pytest.mark.gpu


def wrap_and_call(fn, acts, labels, dtype=torch.float32):
    """Mirror of pytorch_binding/test/test.py:27-48."""
    acts = torch.tensor(acts, dtype=dtype).cuda()
    acts.requires_grad = True
    lengths = torch.IntTensor([acts.shape[1]] * acts.shape[0]).cuda()
    label_lengths = torch.IntTensor([len(l) for l in labels]).cuda()
    labels = torch.IntTensor(labels).cuda()
    costs = fn(acts, labels, lengths, label_lengths)
    cost = torch.sum(costs)
    cost.backward()
    return costs.data.cpu().numpy(), acts.grad.data.cpu().numpy()


def test_small(known_answers):
    from warprnnt_pytorch import RNNTLoss
    ka = known_answers["small"]
    acts = np.array(ka["acts"]).reshape(ka["shape"])
    cost, grads = wrap_and_call(RNNTLoss(reduction='sum'), acts, ka["labels"])
    assert np.allclose(cost, ka["cost"], rtol=1e-6)                       # test.py:75
    assert np.allclose(grads.reshape(-1), ka["logits_grads"], atol=1e-6)   # test.py:77


def test_big(known_answers):
    from warprnnt_pytorch import RNNTLoss
    ka = known_answers["options"]
    acts = np.array(ka["acts_f64"]).reshape(ka["shape"])
    costs, grads = wrap_and_call(RNNTLoss(reduction='sum'), acts, ka["labels"])
    assert np.allclose(costs, sum(ka["costs"]))                            # test.py:155
    assert np.allclose(grads.reshape(-1), ka["logits_grads_hi"], rtol=1e-3)  # test.py:158
    costs, grads = wrap_and_call(RNNTLoss(reduction='sum'), acts, ka["labels"], torch.float64)
    assert np.allclose(costs, sum(ka["costs"]), rtol=1e-12)
    assert np.allclose(grads.reshape(-1), ka["logits_grads_hi"], rtol=1e-6)


def test_reductions_and_grad_output():
    from warprnnt_pytorch import RNNTLoss, rnnt_loss
    rng = np.random.default_rng(1)
    N, T, U, V = 5, 12, 6, 28
    acts_np = rng.standard_normal((N, T, U, V)).astype(np.float32)
    labels_np = rng.integers(1, V, size=(N, U - 1)).astype(np.int32)
    tl_np = np.array([T, 9, 12, 7, 10], np.int32)
    ul_np = np.array([U - 1, 2, 0, 5, 3], np.int32)
    c_ref, g_ref, _ = pyoracle.rnnt_logits(acts_np.astype(np.float64), labels_np, tl_np, ul_np, 0)
    labels, tl, ul = (torch.as_tensor(x).cuda() for x in (labels_np, tl_np, ul_np))

    def run(reduction, weight=None):
        acts = torch.tensor(acts_np, device="cuda", requires_grad=True)
        out = RNNTLoss(reduction=reduction)(acts, labels, tl, ul)
        assert out.is_cuda
        if weight is None:
            out.sum().backward()
        else:
            (out * weight).sum().backward()
        return out.detach().cpu().numpy(), acts.grad.cpu().numpy()

    out, g = run('none')
    assert out.shape == (N,)
    assert np.allclose(out, c_ref, rtol=1e-5) and np.allclose(g, g_ref, rtol=1e-4, atol=1e-6)
    out, g = run('sum')
    assert out.shape == (1,)
    assert np.allclose(out, c_ref.sum(), rtol=1e-5) and np.allclose(g, g_ref, rtol=1e-4, atol=1e-6)
    out, g = run('mean')      # reference: divides by the batch size (:36-40)
    assert np.allclose(out, c_ref.sum() / N, rtol=1e-5)
    assert np.allclose(g, g_ref / N, rtol=1e-4, atol=1e-6)
    w = torch.tensor([1.0, -2.0, 0.5, 3.0, 0.0], device="cuda")
    out, g = run('none', w)   # per-utterance upstream gradient (backward :47-50)
    assert np.allclose(g, g_ref * w.cpu().numpy()[:, None, None, None], rtol=1e-4, atol=1e-6)
    # functional form, default reduction 'mean', no grad required -> loss only
    acts = torch.tensor(acts_np, device="cuda")
    out = rnnt_loss(acts, labels, tl, ul)
    assert np.allclose(out.cpu().numpy(), c_ref.sum() / N, rtol=1e-5)


def test_blank_argument():
    from warprnnt_pytorch import RNNTLoss
    rng = np.random.default_rng(2)
    N, T, U, V = 2, 6, 4, 9
    acts_np = rng.standard_normal((N, T, U, V)).astype(np.float32)
    blank = V - 1
    labels_np = rng.integers(0, V - 1, size=(N, U - 1)).astype(np.int32)
    tl_np, ul_np = np.full(N, T, np.int32), np.full(N, U - 1, np.int32)
    c_ref, g_ref, _ = pyoracle.rnnt_logits(acts_np.astype(np.float64), labels_np, tl_np, ul_np, blank)
    acts = torch.tensor(acts_np, device="cuda", requires_grad=True)
    out = RNNTLoss(blank=blank, reduction='none')(acts, *(torch.as_tensor(x).cuda() for x in (labels_np, tl_np, ul_np)))
    out.sum().backward()
    assert np.allclose(out.detach().cpu().numpy(), c_ref, rtol=1e-5)
    assert np.allclose(acts.grad.cpu().numpy(), g_ref, rtol=1e-4, atol=1e-6)


def test_reference_style_gpu_rnnt_call():
    """The raw extension call the reference's autograd function makes (__init__.py:22-34):
    CPU costs tensor, zero-initialised grads, returns 0."""
    from warprnnt_pytorch import warp_rnnt
    rng = np.random.default_rng(3)
    N, T, U, V = 3, 8, 4, 28
    acts_np = rng.random((N, T, U, V)).astype(np.float32)
    labels_np = rng.integers(1, V, size=(N, U - 1)).astype(np.int32)
    tl_np, ul_np = np.full(N, T, np.int32), np.full(N, U - 1, np.int32)
    acts = torch.tensor(acts_np).cuda()
    grads = torch.zeros_like(acts)
    costs = torch.zeros(N)
    rc = warp_rnnt.gpu_rnnt(acts, *(torch.as_tensor(x).cuda() for x in (labels_np, tl_np, ul_np)),
                            costs, grads, 0, 0)
    assert rc == 0
    c_ref, g_ref, _ = pyoracle.rnnt_logits(acts_np.astype(np.float64), labels_np, tl_np, ul_np, 0)
    assert np.allclose(costs.numpy(), c_ref, rtol=1e-5)
    assert np.allclose(grads.cpu().numpy(), g_ref, rtol=1e-4, atol=1e-6)
    with pytest.raises(RuntimeError):
        warp_rnnt.cpu_rnnt(acts.cpu(), None, None, None, costs, grads.cpu(), 0, 0)


def test_side_stream_retain_graph_and_forward_only():
    """backward twice (retain_graph) accumulates; the op honours the current (non-default) stream;
    no-grad inputs skip the beta lattice but return the same costs."""
    from warprnnt_pytorch import RNNTLoss
    rng = np.random.default_rng(5)
    N, T, U, V = 3, 10, 5, 64
    acts_np = rng.standard_normal((N, T, U, V)).astype(np.float32)
    labels_np = rng.integers(1, V, size=(N, U - 1)).astype(np.int32)
    tl_np = np.array([T, 7, 9], np.int32)
    ul_np = np.array([U - 1, 1, 3], np.int32)
    c_ref, g_ref, _ = pyoracle.rnnt_logits(acts_np.astype(np.float64), labels_np, tl_np, ul_np, 0)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        acts = torch.tensor(acts_np, device="cuda", requires_grad=True)
        labels, tl, ul = (torch.as_tensor(x).cuda() for x in (labels_np, tl_np, ul_np))
        loss = RNNTLoss(reduction='sum')(acts, labels, tl, ul)
        loss.backward(retain_graph=True)
        loss.backward()
        no_grad = RNNTLoss(reduction='none')(acts.detach(), labels, tl, ul)
    side.synchronize()
    assert np.allclose(loss.item(), c_ref.sum(), rtol=1e-5)
    assert np.allclose(acts.grad.cpu().numpy(), 2 * g_ref, rtol=1e-4, atol=2e-6)
    assert np.allclose(no_grad.cpu().numpy(), c_ref, rtol=1e-5)


def test_forward_backward_split_matches_full_call():
    """rnnt_b200_forward + rnnt_b200_backward == compute_rnnt_loss_async, bit for bit at scale 1,
    and per-utterance grad_costs scale rows of the gradient."""
    from warprnnt_pytorch import warp_rnnt as wr
    rng = np.random.default_rng(6)
    N, T, U, V = 4, 9, 6, 300
    acts = torch.tensor(rng.standard_normal((N, T, U, V)).astype(np.float32), device="cuda")
    labels = torch.as_tensor(rng.integers(1, V, size=(N, U - 1)).astype(np.int32)).cuda()
    tl = torch.tensor([T, 5, 9, 3], dtype=torch.int32).cuda()
    ul = torch.tensor([U - 1, 0, 2, 5], dtype=torch.int32).cuda()
    c_full, g_full = torch.empty(N, device="cuda"), torch.empty_like(acts)
    wr.gpu_rnnt_async(acts, labels, tl, ul, c_full, g_full, 0)
    c_split, g_split = torch.empty(N, device="cuda"), torch.full_like(acts, float("nan"))
    ws = wr.gpu_rnnt_forward(acts, labels, tl, ul, c_split, 0, prepare_backward=True)
    wr.gpu_rnnt_backward(acts, labels, tl, ul, g_split, None, 0, 1.0, ws)
    torch.cuda.synchronize()
    assert torch.equal(c_full, c_split) and torch.equal(g_full, g_split)
    w = torch.tensor([2.0, -1.0, 0.0, 0.5], device="cuda")
    wr.gpu_rnnt_backward(acts, labels, tl, ul, g_split, w, 0, 0.25, ws)
    torch.cuda.synchronize()
    assert torch.allclose(g_split, g_full * (0.25 * w).view(-1, 1, 1, 1), rtol=1e-6, atol=0)


def test_async_entry_is_cuda_graph_capturable():
    """compute_rnnt_loss_async makes no allocation, no synchronisation and no host round trip, so a
    training step's loss can be captured into a CUDA graph and replayed on new data."""
    from warprnnt_pytorch import warp_rnnt as wr
    rng = np.random.default_rng(8)
    N, T, U, V = 4, 12, 5, 28
    acts = torch.tensor(rng.standard_normal((N, T, U, V)).astype(np.float32), device="cuda")
    labels_np = rng.integers(1, V, size=(N, U - 1)).astype(np.int32)
    labels = torch.as_tensor(labels_np).cuda()
    tl_np, ul_np = np.full(N, T, np.int32), np.full(N, U - 1, np.int32)
    tl, ul = torch.as_tensor(tl_np).cuda(), torch.as_tensor(ul_np).cuda()
    costs, grads = torch.empty(N, device="cuda"), torch.empty_like(acts)
    ws = torch.empty(wr.workspace_size(T, U, N, 4), dtype=torch.uint8, device="cuda")
    wr.gpu_rnnt_async(acts, labels, tl, ul, costs, grads, 0, 1.0, ws)      # warm-up outside capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        wr.gpu_rnnt_async(acts, labels, tl, ul, costs, grads, 0, 1.0, ws)
    new = rng.standard_normal((N, T, U, V)).astype(np.float32)
    acts.copy_(torch.tensor(new))
    g.replay()
    torch.cuda.synchronize()
    c_ref, g_ref, _ = pyoracle.rnnt_logits(new.astype(np.float64), labels_np, tl_np, ul_np, 0)
    assert np.allclose(costs.cpu().numpy(), c_ref, rtol=1e-5)
    assert np.allclose(grads.cpu().numpy(), g_ref, rtol=1e-4, atol=1e-6)


def test_native_extension_module_matches():
    """warp-transducer_b200/binding/binding.cpp: the compiled (pybind11/libtorch) form of the
    reference's `warp_rnnt` extension module, same call as the reference's _RNNT.forward makes."""
    import importlib.util
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      "warp-transducer_b200", "lib", "warp_rnnt_native.so")
    if not os.path.exists(so):
        pytest.skip("native binding not built")
    spec = importlib.util.spec_from_file_location("warp_rnnt_native", so)
    native = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(native)
    rng = np.random.default_rng(11)
    N, T, U, V = 3, 9, 4, 28
    acts_np = rng.random((N, T, U, V)).astype(np.float32)
    labels_np = rng.integers(1, V, size=(N, U - 1)).astype(np.int32)
    tl_np = np.array([T, 6, 9], np.int32)
    ul_np = np.array([U - 1, 2, 0], np.int32)
    c_ref, g_ref, _ = pyoracle.rnnt_logits(acts_np.astype(np.float64), labels_np, tl_np, ul_np, 0)
    for dt in (torch.float32, torch.float64):
        acts = torch.tensor(acts_np, dtype=dt).cuda()
        grads = torch.zeros_like(acts)
        costs = torch.zeros(N, dtype=dt)
        rc = native.gpu_rnnt(acts, *(torch.as_tensor(x).cuda() for x in (labels_np, tl_np, ul_np)), costs, grads, 0, 0)
        assert rc == 0
        assert np.allclose(costs.numpy(), c_ref, rtol=1e-5)
        assert np.allclose(grads.cpu().numpy(), g_ref, rtol=1e-4, atol=1e-6)
    with pytest.raises(RuntimeError):
        native.cpu_rnnt(acts.cpu(), acts.cpu(), acts.cpu(), acts.cpu(), costs, grads.cpu(), 0, 0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_sixteen_bit_storage(dtype):
    """bf16 / fp16 logits and gradients: fp32 arithmetic inside, so the only error
    against the fp64 oracle evaluated on the SAME rounded inputs is the rounding of the output."""
    from warprnnt_pytorch import RNNTLoss, warp_rnnt as wr
    rng = np.random.default_rng(13)
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    for (N, T, U, V) in [(3, 11, 5, 64), (2, 7, 4, 5000), (2, 6, 3, 37), (2, 9, 34, 16), (2, 5, 3, 4100), (1, 4, 2, 5001)]:
        acts_t = torch.tensor(rng.standard_normal((N, T, U, V)).astype(np.float32)).to(dtype)
        acts_np = acts_t.float().numpy()                      # the rounded logits the kernel sees
        labels_np = rng.integers(1, V, size=(N, U - 1)).astype(np.int32)
        tl_np = rng.integers(T // 2 + 1, T + 1, size=N).astype(np.int32)
        ul_np = rng.integers(0, U, size=N).astype(np.int32)
        tl_np[0], ul_np[0] = T, U - 1
        c_ref, g_ref, _ = pyoracle.rnnt_logits(acts_np.astype(np.float64), labels_np, tl_np, ul_np, 0)
        acts = acts_t.cuda().requires_grad_(True)
        labels, tl, ul = (torch.as_tensor(x).cuda() for x in (labels_np, tl_np, ul_np))
        out = RNNTLoss(reduction='none')(acts, labels, tl, ul)
        assert out.dtype == torch.float32
        out.sum().backward()
        assert acts.grad.dtype == dtype
        assert np.allclose(out.detach().cpu().numpy(), c_ref, rtol=1e-5)
        g = acts.grad.float().cpu().numpy()
        assert np.allclose(g, g_ref, rtol=2 * eps, atol=1e-6), np.abs(g - g_ref).max()
        # full (non-split) entry as well
        c2 = torch.empty(N, device="cuda")
        g2 = torch.empty_like(acts)
        wr.gpu_rnnt_async(acts.detach(), labels, tl, ul, c2, g2, 0)
        torch.cuda.synchronize()
        assert torch.equal(g2, acts.grad) and torch.equal(c2, out.detach())


def test_concurrent_host_threads():
    """The library keeps no shared mutable state (per-thread launch counters / event pools / side
    streams, caller-owned workspace): four host threads on four streams, different shapes."""
    import threading
    from warprnnt_pytorch import warp_rnnt as wr
    shapes = [(3, 11, 5, 28), (2, 9, 40, 6), (2, 6, 4, 1028), (4, 30, 8, 50)]
    results, errors = {}, []

    def work(idx, shape):
        try:
            N, T, U, V = shape
            rng = np.random.default_rng(100 + idx)
            acts_np = rng.standard_normal((N, T, U, V)).astype(np.float32)
            labels_np = rng.integers(1, V, size=(N, U - 1)).astype(np.int32)
            tl_np = rng.integers(T // 2 + 1, T + 1, size=N).astype(np.int32)
            ul_np = rng.integers(0, U, size=N).astype(np.int32)
            c_ref, g_ref, _ = pyoracle.rnnt_logits(acts_np.astype(np.float64), labels_np, tl_np, ul_np, 0, threads=1)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                acts = torch.tensor(acts_np, device="cuda")
                labels, tl, ul = (torch.as_tensor(x).cuda() for x in (labels_np, tl_np, ul_np))
                costs, grads = torch.empty(N, device="cuda"), torch.empty_like(acts)
                for _ in range(25):
                    grads.fill_(float("nan"))
                    wr.gpu_rnnt_async(acts, labels, tl, ul, costs, grads, 0)
                stream.synchronize()
            ok = np.allclose(costs.cpu().numpy(), c_ref, rtol=1e-5) and \
                np.allclose(grads.cpu().numpy(), g_ref, rtol=1e-4, atol=1e-6)
            results[idx] = ok
        except Exception as ex:      # surfaced in the main thread below
            errors.append(repr(ex))

    threads = [threading.Thread(target=work, args=(i, s)) for i, s in enumerate(shapes)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert all(results.get(i) for i in range(len(shapes))), results


def test_length_checks_on_cuda_tensors_raise_and_leave_the_device_usable():
    """The T / U consistency test of certify_inputs is waited for AFTER the kernels are queued (LengthCheck);
    it must still raise from the same call, a labels tensor narrower than U-1 must never reach a kernel,
    and the device must be usable afterwards."""
    from warprnnt_pytorch import RNNTLoss
    rng = np.random.default_rng(3)
    N, T, U, V = 3, 9, 5, 17
    acts = torch.tensor(rng.standard_normal((N, T, U, V)).astype(np.float32), device="cuda", requires_grad=True)
    labels = torch.as_tensor(rng.integers(1, V, size=(N, U - 1)).astype(np.int32)).cuda()
    tl = torch.tensor([T, 7, 5], dtype=torch.int32).cuda()
    ul = torch.tensor([U - 1, 2, 0], dtype=torch.int32).cuda()
    f = RNNTLoss(reduction='sum')
    with pytest.raises(ValueError, match="Input length mismatch"):
        f(acts, labels, torch.tensor([T - 1, 7, 5], dtype=torch.int32).cuda(), ul)
    with pytest.raises(ValueError, match="Input length mismatch"):
        f(acts, labels, torch.tensor([T + 3, 7, 5], dtype=torch.int32).cuda(), ul)     # longer than the tensor: clamped
    with pytest.raises(ValueError, match="Output length mismatch"):
        f(acts, labels, tl, torch.tensor([U - 2, 2, 0], dtype=torch.int32).cuda())
    with pytest.raises(ValueError):
        f(acts, labels[:, :2].contiguous(), tl, ul)                                     # labels narrower than U-1
    loss = f(acts, labels, tl, ul)
    loss.backward()
    torch.cuda.synchronize()
    c_ref, g_ref, _ = pyoracle.rnnt_logits(acts.detach().cpu().numpy().astype(np.float64), labels.cpu().numpy(),
                                           tl.cpu().numpy(), ul.cpu().numpy(), 0)
    assert np.allclose(loss.item(), c_ref.sum(), rtol=1e-5)
    assert np.allclose(acts.grad.cpu().numpy(), g_ref, rtol=1e-4, atol=1e-6)

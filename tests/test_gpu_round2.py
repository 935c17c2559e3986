"""Round-2 parity cases (all through the C-ABI): shared-memory boundaries of the wavefront kernels,
NaN propagation, probabilities far below the fp32 range, the time-major activation layout, and a
tighter element-wise gradient tolerance on large vocabularies."""
import numpy as np
import pytest
import torch

from oracle import pyoracle
from test_gpu_parity import call_abi, check_against, make_inputs, rel_diff, wr  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("U,dtype", [(737, np.float32), (768, np.float32), (769, np.float32),
                                     (353, np.float64), (384, np.float64), (385, np.float64)])
def test_wavefront_shared_memory_boundary(wr, U, dtype):
    """The cp.async ring plus the kernel's static shared memory crosses the 48 KB default exactly in
    these label extents: the opt-in must count both."""
    acts, labels, tl, ul = make_inputs(31, 1, 3, U, 3, dtype=dtype, ragged=False)
    c_ref, g_ref, _ = pyoracle.rnnt_logits(acts.astype(np.float64), labels, tl, ul, 0)
    costs, g = call_abi(wr, acts, labels, tl, ul)
    if dtype == np.float32:
        check_against(costs, g, c_ref, g_ref, "U=%d" % U)
    else:
        assert np.allclose(costs, c_ref, rtol=1e-11) and np.allclose(g, g_ref, rtol=1e-8, atol=1e-13)


@pytest.mark.parametrize("U", [5, 70])
def test_nan_logit_gives_nan_cost(wr, U):
    """A NaN logit must reach the cost of ITS utterance (the reference's log_plus propagates NaN in
    either operand, rnnt_helper.h:16-24) and leave the other utterances alone."""
    acts, labels, tl, ul = make_inputs(8, 3, 9, U, 6, ragged=False)
    clean, _ = call_abi(wr, acts, labels, tl, ul)
    acts[1, 4, 2, 3] = np.nan
    costs, g = call_abi(wr, acts, labels, tl, ul)
    assert np.isnan(costs[1])
    assert np.array_equal(costs[[0, 2]], clean[[0, 2]])
    assert np.isnan(g[1]).any() and np.isfinite(g[0]).all() and np.isfinite(g[2]).all()
    c64, _ = call_abi(wr, acts.astype(np.float64), labels, tl, ul)
    assert np.isnan(c64[1]) and np.isfinite(c64[[0, 2]]).all()


def test_probabilities_far_below_fp32_range(wr):
    """Transition log-probabilities of -1e3 .. -1e4 nats: e^lp underflows fp32 (and fp64) by thousands
    of orders of magnitude, the exponent-carrying wavefront must still agree with the log-domain oracle."""
    for scale, U in ((400.0, 6), (3000.0, 40)):
        acts, labels, tl, ul = make_inputs(13, 2, 14, U, 9, dist="normal")
        acts = (acts * scale / 5.0).astype(np.float32)
        c_ref, g_ref, _ = pyoracle.rnnt_logits(acts.astype(np.float64), labels, tl, ul, 0)
        costs, g = call_abi(wr, acts, labels, tl, ul)
        assert np.all(np.isfinite(costs))
        assert np.allclose(costs, c_ref, rtol=1e-5), (costs, c_ref)
        assert np.allclose(g, g_ref, rtol=1e-3, atol=2e-5), np.abs(g - g_ref).max()


def test_impossible_alignment_is_infinite_cost(wr):
    """-inf on the only path (blank logit of the last frame) -> probability 0 -> cost +inf, as the
    log-domain recurrence gives; the other utterance is unaffected."""
    acts, labels, tl, ul = make_inputs(3, 2, 5, 3, 4, ragged=False)
    clean, _ = call_abi(wr, acts, labels, tl, ul, want_grad=False)
    acts[0, :, :, 0] = -np.inf          # no blank transition is ever possible in utterance 0
    costs, _ = call_abi(wr, acts, labels, tl, ul, want_grad=False)
    assert np.isposinf(costs[0]) and costs[1] == clean[1]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_time_major_layout(wr, dtype):
    """[T,U,N,V] activations through rnnt_b200_loss_async_layout == the batch-first result transposed
    (bit for bit: same arithmetic per row), and == the oracle run on the reference's own
    batch_first=false indexing (cpu_rnnt.h:139-144)."""
    dev = torch.device("cuda:0")
    for (N, T, U, V) in ((3, 7, 4, 28), (5, 6, 3, 50), (2, 4, 3, 1000)):
        acts, labels, tl, ul = make_inputs(17, N, T, U, V, dtype=dtype)
        c0, g0 = call_abi(wr, acts, labels, tl, ul)
        a = torch.as_tensor(np.ascontiguousarray(acts.transpose(1, 2, 0, 3))).to(dev)
        g = torch.full_like(a, float("nan"))
        costs = torch.empty(N, device=dev, dtype=a.dtype)
        lab, tl_d, ul_d = (torch.as_tensor(x).to(dev) for x in (labels, tl, ul))
        wr.gpu_rnnt_async_tunv(a, lab, tl_d, ul_d, costs, g, 0)
        torch.cuda.synchronize()
        assert np.array_equal(costs.cpu().numpy(), c0)
        assert np.array_equal(g.cpu().numpy().transpose(2, 0, 1, 3), g0)
    # the C entry refuses unknown layouts
    st = wr.lib().rnnt_b200_loss_async_layout(7, a.data_ptr(), None, lab.data_ptr(), ul_d.data_ptr(),
                                              tl_d.data_ptr(), V, N, costs.data_ptr(), 1.0, g.data_ptr(),
                                              wr._options(a, 0))
    assert st == 2


def test_batch_first_flag_is_ignored_like_the_reference_gpu_path(wr):
    """tests/test_gpu.cu:42-50 leaves options.batch_first zero-initialised (false) and passes
    [N,T,U,V] data; the reference's GPU kernels never look at the flag (gpu_rnnt_kernel.h:7)."""
    acts, labels, tl, ul = make_inputs(19, 3, 6, 4, 28)
    c0, g0 = call_abi(wr, acts, labels, tl, ul)
    dev = torch.device("cuda:0")
    a = torch.as_tensor(acts).to(dev)
    g = torch.empty_like(a)
    lab, tl_d, ul_d = (torch.as_tensor(x).to(dev) for x in (labels, tl, ul))
    costs = np.zeros(3, np.float32)
    ws = torch.empty(wr.workspace_size(6, 4, 3, 4), dtype=torch.uint8, device=dev)
    opt = wr.rnntOptions(loc=1, num_threads=0, stream=None, blank_label=0, maxT=6, maxU=4, batch_first=False)
    st = wr.lib().compute_rnnt_loss(a.data_ptr(), g.data_ptr(), lab.data_ptr(), ul_d.data_ptr(), tl_d.data_ptr(),
                                    28, 3, costs.ctypes.data, ws.data_ptr(), opt)
    assert st == 0 and np.array_equal(costs, c0) and np.array_equal(g.cpu().numpy(), g0)


@pytest.mark.parametrize("shape", [(2, 6, 4, 1000), (2, 5, 3, 5000), (1, 4, 3, 8200)])
def test_large_vocab_gradient_tolerance_1e7(wr, shape):
    """at V >= 1000 a typical gradient element is ~1e-4, so the element-wise floor
    is 1e-7 (not 1e-6): |g - g_ref| <= 1e-4 |g_ref| + 1e-7."""
    N, T, U, V = shape
    acts, labels, tl, ul = make_inputs(23, N, T, U, V)
    c_ref, g_ref, _ = pyoracle.rnnt_logits(acts.astype(np.float64), labels, tl, ul, 0)
    costs, g = call_abi(wr, acts, labels, tl, ul)
    assert np.allclose(costs, c_ref, rtol=1e-6)
    bad = ~np.isclose(g, g_ref, rtol=1e-4, atol=1e-7)
    assert not bad.any(), (int(bad.sum()), np.abs(g - g_ref).max())
    assert rel_diff(g, g_ref) < 1e-10


def test_forward_backward_likelihoods_agree(wr):
    """The reference's debug guard (cpu_rnnt.h:167-170): |llForward - llBackward| small.  The backward
    likelihood is in the workspace after every gradient call; read it back through the test hook."""
    acts, labels, tl, ul = make_inputs(29, 4, 30, 9, 28)
    dev = torch.device("cuda:0")
    a = torch.as_tensor(acts).to(dev)
    g = torch.empty_like(a)
    costs = torch.empty(4, device=dev)
    lab, tl_d, ul_d = (torch.as_tensor(x).to(dev) for x in (labels, tl, ul))
    ws = wr.gpu_rnnt_async(a, lab, tl_d, ul_d, costs, g, 0)
    llf, llb = wr.read_log_likelihoods(ws, 30, 9, 4, 4)
    assert np.allclose(llf, llb, rtol=1e-6, atol=1e-4), (llf, llb)
    assert np.allclose(-llf, costs.cpu().numpy(), rtol=1e-6)


def test_tensorflow_op_call_sequence(wr, known_answers):
    """The TensorFlow GPU kernel's exact calls (tensorflow_binding/src/warprnnt_op.cc:97-129,173-187):
    options = rnntOptions{} with only loc / stream / blank_label / maxT / maxU set (batch_first stays
    false), get_workspace_size with the default dtype size, device acts / grads / labels / lengths, and
    `costs` in HOST memory (.HostMemory("costs")), readable when the call returns.  Checked against the
    vectors of tensorflow_binding/tests/test_warprnnt_op.py:68-79."""
    import ctypes as C
    ka = known_answers["options"]
    a = np.array(ka["acts"], np.float32).reshape(ka["shape"])
    N, T, U, V = a.shape
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        acts = torch.as_tensor(a).to(dev)
        grads = torch.full_like(acts, float("nan"))
        labels = torch.as_tensor(np.array(ka["labels"], np.int32)).to(dev)
        tl = torch.full((N,), T, dtype=torch.int32, device=dev)
        ul = torch.full((N,), U - 1, dtype=torch.int32, device=dev)
    stream.synchronize()
    opt = wr.rnntOptions()                      # zero-initialised, as `auto options = rnntOptions{}`
    opt.loc, opt.stream, opt.blank_label, opt.maxT, opt.maxU = 1, stream.cuda_stream, 0, T, U
    nbytes = C.c_size_t(0)
    assert wr.lib().get_workspace_size(T, U, N, True, C.byref(nbytes), 4) == 0
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)      # ctx->allocate_temp
    costs = np.full(N, np.nan, np.float32)                             # host memory
    st = wr.lib().compute_rnnt_loss(acts.data_ptr(), grads.data_ptr(), labels.data_ptr(), ul.data_ptr(),
                                    tl.data_ptr(), V, N, costs.ctypes.data, ws.data_ptr(), opt)
    assert st == 0
    assert np.allclose(costs, ka["costs"], atol=1e-6)                  # readable right after the call
    assert np.allclose(grads.cpu().numpy().reshape(-1), ka["logits_grads"], atol=1e-6)   # the TF test's tolerance


def test_every_gradient_element_is_written():
    """The library writes EVERY element of the gradient tensors (zeros on padding, no memset pass):
    buffers pre-filled with NaN must come back NaN-free - every kernel family, ragged lengths, the
    training-step split, 16-bit storage, fp64, and both additive-joint gradient paths."""
    import torch
    from warprnnt_pytorch import warp_rnnt as w
    from warprnnt_pytorch.joint import add_joint_call
    rng = np.random.default_rng(5)
    dev = torch.device("cuda:0")
    shapes = [(3, 9, 5, 28), (2, 12, 40, 6), (2, 7, 4, 301), (2, 5, 3, 1028), (1, 3, 2, 8200), (9, 30, 70, 50),
              (4, 33, 9, 50), (3, 10, 3, 64), (2, 6, 4, 5000), (2, 3, 129, 5)]
    for dt in (torch.float32, torch.float64, torch.bfloat16):
        for (N, T, U, V) in shapes:
            if dt != torch.float32 and V > 2000 and U > 4:
                continue
            acts = torch.tensor(rng.standard_normal((N, T, U, V)).astype(np.float32), device=dev).to(dt)
            labels = torch.as_tensor(rng.integers(1, V, size=(N, U - 1)).astype(np.int32)).to(dev)
            tl = torch.as_tensor(rng.integers(T // 2 + 1, T + 1, size=N).astype(np.int32)).to(dev)
            ul = torch.as_tensor(rng.integers(0, U, size=N).astype(np.int32)).to(dev)
            costs = torch.empty(N, device=dev, dtype=torch.float64 if dt == torch.float64 else torch.float32)
            grads = torch.full_like(acts, float("nan"))
            w.gpu_rnnt_async(acts, labels, tl, ul, costs, grads, 0)
            torch.cuda.synchronize()
            assert not torch.isnan(grads.float()).any(), ("full", dt, (N, T, U, V))
            grads = torch.full_like(acts, float("nan"))
            ws = w.gpu_rnnt_forward(acts, labels, tl, ul, costs, 0, True)
            w.gpu_rnnt_backward(acts, labels, tl, ul, grads, None, 0, 0.5, ws)
            torch.cuda.synchronize()
            assert not torch.isnan(grads.float()).any(), ("split", dt, (N, T, U, V))
    for (N, T, U, V) in [(2, 9, 5, 28), (2, 20, 34, 700), (2, 40, 7, 130), (2, 70, 21, 520), (3, 33, 32, 64)]:
        trans = torch.tensor(rng.standard_normal((N, T, V)).astype(np.float32), device=dev)
        pred = torch.tensor(rng.standard_normal((N, U, V)).astype(np.float32), device=dev)
        labels = torch.as_tensor(rng.integers(1, V, size=(N, U - 1)).astype(np.int32)).to(dev)
        tl = torch.as_tensor(rng.integers(T // 2 + 1, T + 1, size=N).astype(np.int32)).to(dev)
        ul = torch.as_tensor(rng.integers(0, U, size=N).astype(np.int32)).to(dev)
        costs = torch.empty(N, device=dev)
        dtrans, dpred = torch.full_like(trans, float("nan")), torch.full_like(pred, float("nan"))
        ws = add_joint_call(trans, pred, labels, tl, ul, costs, dtrans, dpred, 0, 1.0)
        torch.cuda.synchronize()
        assert not torch.isnan(dtrans).any() and not torch.isnan(dpred).any(), ("joint", (N, T, U, V))
        del ws

"""`warprnnt_pytorch.warp_rnnt` — the extension-module surface of the reference binding
(pytorch_binding/src/binding.cpp:12-19,84-91,157-162), bound to libwarprnnt.so's C-ABI.

    gpu_rnnt(acts, labels, input_lengths, label_lengths, costs, grads, blank_label, num_threads) -> int
    cpu_rnnt(...)   raises: this build has no CPU path (and never falls back to one)

The library is loaded eagerly; a missing or unloadable libwarprnnt.so is an ImportError, not a
silent fallback.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.environ.get("WARP_RNNT_LIB", os.path.join(_PKG, "lib", "libwarprnnt.so"))


class rnntOptions(C.Structure):
    """include/rnnt.h `struct rnntOptions` (32 bytes, by value)."""
    _fields_ = [("loc", C.c_int), ("num_threads", C.c_uint), ("stream", C.c_void_p),
                ("blank_label", C.c_int), ("maxT", C.c_int), ("maxU", C.c_int),
                ("batch_first", C.c_bool)]


RNNT_CPU, RNNT_GPU = 0, 1
RNNT_STATUS_SUCCESS = 0

if not os.path.exists(_LIB_PATH):
    raise ImportError(
        "libwarprnnt.so not found at %s - build it with `python warp-transducer_b200/build.py` "
        "(there is no CPU or PyTorch fallback)" % _LIB_PATH)
_lib = C.CDLL(_LIB_PATH)
assert C.sizeof(rnntOptions) == 32

_P = C.c_void_p
for _name in ("compute_rnnt_loss", "compute_rnnt_loss_fp64"):
    _f = getattr(_lib, _name)
    _f.restype = C.c_int
    _f.argtypes = [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, rnntOptions]
_lib.compute_rnnt_loss_async.restype = C.c_int
_lib.compute_rnnt_loss_async.argtypes = [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_float, _P, rnntOptions]
_lib.compute_rnnt_loss_async_fp64.restype = C.c_int
_lib.compute_rnnt_loss_async_fp64.argtypes = [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_double, _P, rnntOptions]
for _name, _ct in (("rnnt_b200_forward", C.c_float), ("rnnt_b200_forward_fp64", C.c_double)):
    _f = getattr(_lib, _name)
    _f.restype = C.c_int
    _f.argtypes = [_P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_int, _P, rnntOptions]
for _name, _ct in (("rnnt_b200_backward", C.c_float), ("rnnt_b200_backward_fp64", C.c_double)):
    _f = getattr(_lib, _name)
    _f.restype = C.c_int
    _f.argtypes = [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _ct, _P, rnntOptions]
_lib.rnnt_b200_loss_async_16.restype = C.c_int
_lib.rnnt_b200_loss_async_16.argtypes = [C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_float, _P, rnntOptions]
_lib.rnnt_b200_forward_16.restype = C.c_int
_lib.rnnt_b200_forward_16.argtypes = [C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_int, _P, rnntOptions]
_lib.rnnt_b200_backward_16.restype = C.c_int
_lib.rnnt_b200_backward_16.argtypes = [C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_float, _P, rnntOptions]
RNNT_B200_BF16, RNNT_B200_FP16 = 1, 2
_lib.rnnt_b200_loss_async_layout.restype = C.c_int
_lib.rnnt_b200_loss_async_layout.argtypes = [C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_float, _P, rnntOptions]
_lib.rnnt_b200_loss_async_layout_fp64.restype = C.c_int
_lib.rnnt_b200_loss_async_layout_fp64.argtypes = [C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_double, _P, rnntOptions]
RNNT_B200_LAYOUT_NTUV, RNNT_B200_LAYOUT_TUNV = 0, 1
_lib.get_workspace_size.restype = C.c_int
_lib.get_workspace_size.argtypes = [C.c_int, C.c_int, C.c_int, C.c_bool, C.POINTER(C.c_size_t), C.c_size_t]
_lib.get_warprnnt_version.restype = C.c_int
_lib.rnntGetStatusString.restype = C.c_char_p
_lib.rnntGetStatusString.argtypes = [C.c_int]
_lib.rnnt_b200_last_launch_count.restype = C.c_int
_lib.rnnt_b200_build_info.restype = C.c_char_p
_lib.rnnt_b200_set_profiling.restype = None
_lib.rnnt_b200_set_profiling.argtypes = [C.c_int]
_lib.rnnt_b200_last_kernel_ms.restype = C.c_int
_lib.rnnt_b200_last_kernel_ms.argtypes = [C.POINTER(C.c_float)]
_lib.rnnt_b200_profile_collect.restype = C.c_int
_lib.rnnt_b200_profile_collect.argtypes = [C.POINTER(C.c_float)]


def lib():
    """The loaded ctypes handle (tests and bench.py call the C-ABI through it)."""
    return _lib


def lib_path():
    return _LIB_PATH


def status_string(status):
    return _lib.rnntGetStatusString(int(status)).decode()


def workspace_size(maxT, maxU, minibatch, dtype_size=4, gpu=True):
    n = C.c_size_t(0)
    st = _lib.get_workspace_size(maxT, maxU, minibatch, gpu, C.byref(n), dtype_size)
    if st != RNNT_STATUS_SUCCESS:
        raise ValueError("get_workspace_size: " + status_string(st))
    return n.value


def last_launch_count():
    return _lib.rnnt_b200_last_launch_count()


def set_profiling(enabled):
    _lib.rnnt_b200_set_profiling(1 if enabled else 0)


def last_kernel_ms():
    """(rowstats, lattice, grad) milliseconds of the last profiled call on this thread."""
    out = (C.c_float * 3)()
    _lib.rnnt_b200_last_kernel_ms(out)
    return tuple(out)


def _options(acts, blank_label, num_threads=0):
    opt = rnntOptions()
    opt.loc = RNNT_GPU
    opt.num_threads = num_threads
    opt.stream = torch.cuda.current_stream(acts.device).cuda_stream
    opt.blank_label = blank_label
    opt.maxT = acts.size(1)
    opt.maxU = acts.size(2)
    opt.batch_first = True
    return opt


def _ptr(t):
    return t.data_ptr() if t is not None and t.numel() > 0 else None


_DUMMY_LABELS = {}   # device -> 1-element int32 tensor, kept alive for the life of the process


def _labels_ptr(labels):
    """U == 1 (no labels at all): the ABI still wants a non-null pointer.  The stand-in is a cached
    per-device tensor - a temporary would be freed before the (asynchronous) launch reads it."""
    if labels.numel() > 0:
        return labels.data_ptr()
    key = (labels.device.type, labels.device.index)
    if key not in _DUMMY_LABELS:
        _DUMMY_LABELS[key] = torch.zeros(1, dtype=torch.int32, device=labels.device)
    return _DUMMY_LABELS[key].data_ptr()


def require_same_device(ref, **tensors):
    """The async entry points dereference every pointer on `ref`'s device (host staging exists only in
    the synchronous C API), so a CPU or other-GPU tensor would be an illegal access, not an error."""
    for name, t in tensors.items():
        if t is not None and t.device != ref.device:
            raise RuntimeError("%s is on %s but the activations are on %s: all operator inputs must "
                               "live on the activations' CUDA device" % (name, t.device, ref.device))


def gpu_rnnt(acts, labels, input_lengths, label_lengths, costs, grads, blank_label, num_threads):
    """Reference signature (binding.cpp:84-91).  `costs` is a CPU tensor [N] (as the reference
    binding requires) or a CUDA tensor; `grads` is like `acts`, or empty for loss only.
    Returns 0 on success, raises on a library error (the reference ignored the status)."""
    if not acts.is_cuda:
        raise RuntimeError("gpu_rnnt needs CUDA tensors")
    N, T, U, V = acts.shape
    if acts.dtype == torch.float32:
        fn, esz = _lib.compute_rnnt_loss, 4
    elif acts.dtype == torch.float64:
        fn, esz = _lib.compute_rnnt_loss_fp64, 8
    else:
        raise TypeError("unsupported data type %s" % acts.dtype)
    with torch.cuda.device(acts.device):
        ws = torch.empty(workspace_size(T, U, N, esz), dtype=torch.uint8, device=acts.device)
        st = fn(acts.data_ptr(), _ptr(grads), _labels_ptr(labels), label_lengths.data_ptr(),
                input_lengths.data_ptr(), V, N, costs.data_ptr(), ws.data_ptr(),
                _options(acts, blank_label, num_threads))
    if st != RNNT_STATUS_SUCCESS:
        raise RuntimeError("compute_rnnt_loss failed: " + status_string(st))
    return 0


def gpu_rnnt_async(acts, labels, input_lengths, label_lengths, costs, grads, blank_label,
                   grad_scale=1.0, workspace=None):
    """Extension: no host synchronisation, `costs` on the device, gradients pre-multiplied by
    `grad_scale`.  Returns the workspace tensor (keep it alive until the stream has run)."""
    N, T, U, V = acts.shape
    code = _code16(acts)
    if code:
        esz = 4
    elif acts.dtype == torch.float32:
        fn, esz = _lib.compute_rnnt_loss_async, 4
    elif acts.dtype == torch.float64:
        fn, esz = _lib.compute_rnnt_loss_async_fp64, 8
    else:
        raise TypeError("unsupported data type %s" % acts.dtype)
    with torch.cuda.device(acts.device):
        need = workspace_size(T, U, N, esz)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=acts.device)
        args = (acts.data_ptr(), _ptr(grads), _labels_ptr(labels), label_lengths.data_ptr(),
                input_lengths.data_ptr(), V, N, costs.data_ptr(), grad_scale, workspace.data_ptr(),
                _options(acts, blank_label))
        st = _lib.rnnt_b200_loss_async_16(code, *args) if code else fn(*args)
    if st != RNNT_STATUS_SUCCESS:
        raise RuntimeError("compute_rnnt_loss_async failed: " + status_string(st))
    return workspace


def gpu_rnnt_async_tunv(acts, labels, input_lengths, label_lengths, costs, grads, blank_label,
                        grad_scale=1.0, workspace=None):
    """Time-major extension: `acts` / `grads` are [T, U, N, V] (the layout the reference's CPU path
    indexes for batch_first == false, cpu_rnnt.h:139-144); labels [N, U-1], lengths and costs [N].
    fp32 / fp64, no host synchronisation.  Returns the workspace tensor."""
    T, U, N, V = acts.shape
    fn, esz = _pick(acts, "rnnt_b200_loss_async_layout", "rnnt_b200_loss_async_layout_fp64")
    with torch.cuda.device(acts.device):
        need = workspace_size(T, U, N, esz)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=acts.device)
        opt = _options(acts, blank_label)
        opt.maxT, opt.maxU = T, U
        st = fn(RNNT_B200_LAYOUT_TUNV, acts.data_ptr(), _ptr(grads), _labels_ptr(labels),
                label_lengths.data_ptr(), input_lengths.data_ptr(), V, N, costs.data_ptr(), grad_scale,
                workspace.data_ptr(), opt)
    if st != RNNT_STATUS_SUCCESS:
        raise RuntimeError("rnnt_b200_loss_async_layout failed: " + status_string(st))
    return workspace


def _code16(acts):
    return {torch.bfloat16: RNNT_B200_BF16, torch.float16: RNNT_B200_FP16}.get(acts.dtype)


def costs_dtype(acts):
    """dtype of the per-utterance costs for a given activation dtype (fp32 for 16-bit storage)."""
    return torch.float32 if _code16(acts) else acts.dtype


def _pick(acts, f32, f64):
    if acts.dtype == torch.float32:
        return getattr(_lib, f32), 4
    if acts.dtype == torch.float64:
        return getattr(_lib, f64), 8
    raise TypeError("unsupported data type %s" % acts.dtype)


def gpu_rnnt_forward(acts, labels, input_lengths, label_lengths, costs, blank_label,
                     prepare_backward=True, workspace=None):
    """Training-step split, first half: statistics + lattices into `workspace`, costs on the
    device, no synchronisation.  Returns the workspace tensor (hand it to gpu_rnnt_backward)."""
    N, T, U, V = acts.shape
    code = _code16(acts)
    fn, esz = (None, 4) if code else _pick(acts, "rnnt_b200_forward", "rnnt_b200_forward_fp64")
    with torch.cuda.device(acts.device):
        need = workspace_size(T, U, N, esz)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=acts.device)
        args = (acts.data_ptr(), _labels_ptr(labels), label_lengths.data_ptr(), input_lengths.data_ptr(),
                V, N, costs.data_ptr(), 1 if prepare_backward else 0, workspace.data_ptr(),
                _options(acts, blank_label))
        st = _lib.rnnt_b200_forward_16(code, *args) if code else fn(*args)
    if st != RNNT_STATUS_SUCCESS:
        raise RuntimeError("rnnt_b200_forward failed: " + status_string(st))
    return workspace


def gpu_rnnt_backward(acts, labels, input_lengths, label_lengths, grads, grad_costs, blank_label,
                      grad_scale, workspace):
    """Second half: grads[b] = grad_scale * grad_costs[b] * d cost[b] / d acts[b] from the lattices
    gpu_rnnt_forward left in `workspace` (grad_costs: device tensor [N] or None for ones)."""
    N, T, U, V = acts.shape
    code = _code16(acts)
    fn = None if code else _pick(acts, "rnnt_b200_backward", "rnnt_b200_backward_fp64")[0]
    with torch.cuda.device(acts.device):
        args = (acts.data_ptr(), grads.data_ptr(), _labels_ptr(labels), label_lengths.data_ptr(),
                input_lengths.data_ptr(), V, N, _ptr(grad_costs), grad_scale, workspace.data_ptr(),
                _options(acts, blank_label))
        st = _lib.rnnt_b200_backward_16(code, *args) if code else fn(*args)
    if st != RNNT_STATUS_SUCCESS:
        raise RuntimeError("rnnt_b200_backward failed: " + status_string(st))
    return 0


_lib.rnnt_b200_debug_log_likelihoods.restype = C.c_int
_lib.rnnt_b200_debug_log_likelihoods.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_size_t, _P, _P]


def read_log_likelihoods(workspace, maxT, maxU, minibatch, dtype_size=4):
    """(llForward, llBackward) of the last loss+gradient call that used `workspace` (test hook)."""
    import numpy as np
    f, b = np.zeros(minibatch), np.zeros(minibatch)
    st = _lib.rnnt_b200_debug_log_likelihoods(workspace.data_ptr(), maxT, maxU, minibatch, dtype_size,
                                              f.ctypes.data, b.ctypes.data)
    if st != RNNT_STATUS_SUCCESS:
        raise RuntimeError("rnnt_b200_debug_log_likelihoods: " + status_string(st))
    return f, b


def profile_collect():
    """(calls, (rowstats, lattice, grad) mean ms) over the profiled calls since the last collect."""
    out = (C.c_float * 3)()
    n = _lib.rnnt_b200_profile_collect(out)
    return n, tuple(out)


def cpu_rnnt(acts, labels, input_lengths, label_lengths, costs, grads, blank_label, num_threads):
    """Reference signature (binding.cpp:12-19).  Not available: the B200 build is the device
    path only and must not fall back to a host implementation."""
    raise RuntimeError("warprnnt_pytorch (B200 build): cpu_rnnt is not available - "
                       "move the tensors to a CUDA device")

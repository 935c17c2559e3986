"""RNN-T loss for an ADDITIVE joint network without materialising the [N,T,U,V] logits
(SURVEY.md §8(f).2).  The reference's own timing script forms its logits as
``acts = trans.unsqueeze(2) + pred.unsqueeze(1)`` (pytorch_binding/test/test_time.py:73) and then
calls RNNTLoss on the 4-D tensor; this operator takes the two factors directly:

    loss = AddJointRNNTLoss(blank=0, reduction='mean')(trans, pred, labels, act_lens, label_lens)

trans [N,T,V], pred [N,U,V] fp32 CUDA, same label/length conventions and input checks as RNNTLoss.
Equivalent to RNNTLoss()(trans.unsqueeze(2) + pred.unsqueeze(1), ...) with gradients reduced to
the factors, at O(N (T+U) V) memory traffic.
"""
import ctypes as C

import torch
from torch.autograd import Function
from torch.nn import Module

from . import warp_rnnt
from ._checks import LengthCheck, check_contiguous, check_dim, check_type

_lib = warp_rnnt.lib()
_P = C.c_void_p
_lib.rnnt_b200_add_joint_loss.restype = C.c_int
_lib.rnnt_b200_add_joint_loss.argtypes = [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_float, _P,
                                          warp_rnnt.rnntOptions]
_lib.rnnt_b200_add_joint_forward.restype = C.c_int
_lib.rnnt_b200_add_joint_forward.argtypes = [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_int, _P,
                                             warp_rnnt.rnntOptions]
_lib.rnnt_b200_add_joint_backward.restype = C.c_int
_lib.rnnt_b200_add_joint_backward.argtypes = [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_float, _P,
                                              warp_rnnt.rnntOptions]
_lib.rnnt_b200_add_joint_workspace_size.restype = C.c_int
_lib.rnnt_b200_add_joint_workspace_size.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]


def certify_joint_inputs(trans, pred, labels, lengths, label_lengths, defer=False):
    check_type(labels, torch.int32, "labels")
    check_type(label_lengths, torch.int32, "label_lengths")
    check_type(lengths, torch.int32, "lengths")
    for t, n in ((trans, "trans"), (pred, "pred"), (labels, "labels"), (lengths, "lengths"),
                 (label_lengths, "label_lengths")):
        check_contiguous(t, n)
    check_dim(trans, 3, "trans")
    check_dim(pred, 3, "pred")
    check_dim(labels, 2, "labels")
    if trans.dtype != torch.float32 or pred.dtype != torch.float32:
        raise TypeError("trans and pred must be float32")
    if not (trans.shape[0] == pred.shape[0] == lengths.shape[0] == label_lengths.shape[0]):
        raise ValueError("must have a length and a label length per example.")
    if trans.shape[2] != pred.shape[2]:
        raise ValueError("trans and pred must share the vocabulary dimension")
    check = LengthCheck(lengths, label_lengths, trans.shape[1], pred.shape[1])
    if defer:
        return check
    check.finish()
    return None


def add_joint_call(trans, pred, labels, act_lens, label_lens, costs, dtrans, dpred, blank, scale):
    """Raw call of rnnt_b200_add_joint_loss on CUDA tensors (no checks, no synchronisation)."""
    N, T, V = trans.shape
    U = pred.shape[1]
    n = C.c_size_t(0)
    st = _lib.rnnt_b200_add_joint_workspace_size(T, U, N, V, C.byref(n))
    if st != 0:
        raise ValueError("workspace size: " + warp_rnnt.status_string(st))
    with torch.cuda.device(trans.device):
        ws = torch.empty(n.value, dtype=torch.uint8, device=trans.device)
        opt = warp_rnnt.rnntOptions(loc=1, num_threads=0,
                                    stream=torch.cuda.current_stream(trans.device).cuda_stream,
                                    blank_label=blank, maxT=T, maxU=U, batch_first=True)
        lab = warp_rnnt._labels_ptr(labels)
        st = _lib.rnnt_b200_add_joint_loss(trans.data_ptr(), pred.data_ptr(),
                                           dtrans.data_ptr() if dtrans is not None else None,
                                           dpred.data_ptr() if dpred is not None else None,
                                           lab, label_lens.data_ptr(), act_lens.data_ptr(), V, N,
                                           costs.data_ptr(), scale, ws.data_ptr(), opt)
    if st != 0:
        raise RuntimeError("rnnt_b200_add_joint_loss failed: " + warp_rnnt.status_string(st))
    return ws


def _joint_opts(trans, pred, blank):
    return warp_rnnt.rnntOptions(loc=1, num_threads=0,
                                 stream=torch.cuda.current_stream(trans.device).cuda_stream,
                                 blank_label=blank, maxT=trans.shape[1], maxU=pred.shape[1], batch_first=True)


_lab_ptr = warp_rnnt._labels_ptr   # cached per-device stand-in when there are no labels (U == 1)


class _AddJointRNNT(Function):
    """forward: factor exponentials, S = Ef.Eg^T, alpha/beta lattices, costs.  backward: weights +
    the two factor-gradient contractions with grad_output[b] and the reduction factor folded in."""

    @staticmethod
    def forward(ctx, trans, pred, labels, act_lens, label_lens, blank, reduction):
        length_check = certify_joint_inputs(trans, pred, labels, act_lens, label_lens, defer=True)
        if not trans.is_cuda:
            raise RuntimeError("warprnnt_pytorch (B200 build) runs on CUDA tensors only")
        warp_rnnt.require_same_device(trans, pred=pred, labels=labels, act_lens=act_lens, label_lens=label_lens)
        if reduction not in ('none', 'sum', 'mean'):
            raise ValueError("reduction must be 'none', 'sum' or 'mean'")
        N, T, V = trans.shape
        U = pred.shape[1]
        length_check.guard_labels(labels, N)
        need = trans.requires_grad or pred.requires_grad
        costs = torch.empty(N, dtype=torch.float32, device=trans.device)
        n = C.c_size_t(0)
        _lib.rnnt_b200_add_joint_workspace_size(T, U, N, V, C.byref(n))
        with torch.cuda.device(trans.device):
            ws = torch.empty(n.value, dtype=torch.uint8, device=trans.device)
            st = _lib.rnnt_b200_add_joint_forward(trans.data_ptr(), pred.data_ptr(), _lab_ptr(labels),
                                                  label_lens.data_ptr(), act_lens.data_ptr(), V, N,
                                                  costs.data_ptr(), 1 if need else 0, ws.data_ptr(),
                                                  _joint_opts(trans, pred, blank))
        if st != 0:
            raise RuntimeError("rnnt_b200_add_joint_forward failed: " + warp_rnnt.status_string(st))
        length_check.finish()   # the reference's length test, waited for with the kernels already queued
        if need:
            ctx.save_for_backward(trans, pred, labels, act_lens, label_lens)
            ctx.ws, ctx.blank = ws, blank
            ctx.scale = 1.0 / N if reduction == 'mean' else 1.0
        if reduction in ('sum', 'mean'):
            costs = costs.sum().unsqueeze_(-1)
            if reduction == 'mean':
                costs /= N
        return costs

    @staticmethod
    def backward(ctx, grad_output):
        trans, pred, labels, act_lens, label_lens = ctx.saved_tensors
        N, T, V = trans.shape
        g = grad_output.reshape(-1).to(device=trans.device, dtype=torch.float32)
        g = g.expand(N).contiguous() if g.numel() == 1 else g.contiguous()
        dtrans, dpred = torch.empty_like(trans), torch.empty_like(pred)
        with torch.cuda.device(trans.device):
            st = _lib.rnnt_b200_add_joint_backward(trans.data_ptr(), pred.data_ptr(), dtrans.data_ptr(),
                                                   dpred.data_ptr(), _lab_ptr(labels), label_lens.data_ptr(),
                                                   act_lens.data_ptr(), V, N, g.data_ptr(), ctx.scale,
                                                   ctx.ws.data_ptr(), _joint_opts(trans, pred, ctx.blank))
        if st != 0:
            raise RuntimeError("rnnt_b200_add_joint_backward failed: " + warp_rnnt.status_string(st))
        return dtrans, dpred, None, None, None, None, None


def add_joint_rnnt_loss(trans, pred, labels, act_lens, label_lens, blank=0, reduction='mean'):
    return _AddJointRNNT.apply(trans, pred, labels, act_lens, label_lens, blank, reduction)


class AddJointRNNTLoss(Module):
    def __init__(self, blank=0, reduction='mean'):
        super().__init__()
        self.blank, self.reduction = blank, reduction

    def forward(self, trans, pred, labels, act_lens, label_lens):
        return _AddJointRNNT.apply(trans, pred, labels, act_lens, label_lens, self.blank, self.reduction)

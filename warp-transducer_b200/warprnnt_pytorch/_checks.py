"""Input validation for the RNN-T operators.

Behavioural contract = the reference's `certify_inputs`
(pytorch_binding/warprnnt_pytorch/__init__.py:115-140): int32 labels / lengths (TypeError),
contiguous tensors, 4-D activations, 2-D labels, 1-D lengths, one length per utterance,
T == max(lengths), U == max(label_lengths) + 1 (all ValueError).  Written table-driven; the
two maxima come back in ONE device-to-host transfer (the reference pays two).
"""
import torch


def check_type(var, t, name):
    if var.dtype is not t:
        raise TypeError("{} must be {}".format(name, t))


def check_contiguous(var, name):
    if not var.is_contiguous():
        raise ValueError("{} must be contiguous".format(name))


def check_dim(var, dim, name):
    if var.dim() != dim:
        raise ValueError("{} must be {}D".format(name, dim))


def certify_inputs(log_probs, labels, lengths, label_lengths):
    named = (("log_probs", log_probs, None, 4), ("labels", labels, torch.int32, 2),
             ("lengths", lengths, torch.int32, 1), ("label_lengths", label_lengths, torch.int32, 1))
    for name, tensor, dtype, _ in named:          # dtypes first, as the reference does
        if dtype is not None:
            check_type(tensor, dtype, name)
    for name, tensor, _, _ in named:
        check_contiguous(tensor, name)
    batch = log_probs.shape[0]
    if lengths.shape[0] != batch:
        raise ValueError("must have a length per example.")
    if label_lengths.shape[0] != batch:
        raise ValueError("must have a label length per example.")
    for name, tensor, _, rank in named:
        check_dim(tensor, rank, name)
    longest = torch.stack((lengths.max(), label_lengths.max())).tolist()
    if log_probs.shape[1] != longest[0]:
        raise ValueError("Input length mismatch")
    if log_probs.shape[2] != longest[1] + 1:
        raise ValueError("Output length mismatch")

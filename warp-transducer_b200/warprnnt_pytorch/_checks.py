"""Input validation for the RNN-T operators.

Behavioural contract = the reference's `certify_inputs`
(pytorch_binding/warprnnt_pytorch/__init__.py:115-140): int32 labels / lengths (TypeError),
contiguous tensors, 4-D activations, 2-D labels, 1-D lengths, one length per utterance,
T == max(lengths), U == max(label_lengths) + 1 (all ValueError).  Written table-driven; the
two maxima come back in ONE device-to-host transfer (the reference pays two), and the operators
wait for it only after their kernels are queued (LengthCheck).
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


class LengthCheck:
    """The reference's `T == max(lengths)`, `U == max(label_lengths) + 1` test without draining the GPU.

    The two maxima are copied to pinned host memory asynchronously BEFORE the operator queues its kernels;
    finish() waits for that copy only (an event recorded right behind it).  By then the operator's kernels are
    queued, so the device never idles while the host compares two integers (a plain .tolist() here made
    every training step start on an empty GPU: ~10 % of the additive-joint step).  The kernels clamp lengths
    into the tensor extents, so having queued them on inconsistent lengths is harmless; the ValueError is
    raised from the same forward() call all the same."""
    _pinned = {}

    def __init__(self, lengths, label_lengths, T, U):
        self.T, self.U, self.event = T, U, None
        if not (lengths.is_cuda and label_lengths.is_cuda):
            self.values = torch.stack((lengths.max(), label_lengths.max())).tolist()
            self.finish()          # host tensors: nothing to overlap, fail right away as the reference does
            return
        import threading
        key = (lengths.device.index, threading.get_ident())
        buf = LengthCheck._pinned.get(key)
        if buf is None:
            buf = LengthCheck._pinned[key] = torch.empty(2, dtype=torch.int32).pin_memory()
        with torch.cuda.device(lengths.device):
            buf.copy_(torch.stack((lengths.max(), label_lengths.max())), non_blocking=True)
            self.event = torch.cuda.Event()
            self.event.record()
        self.values = buf

    def guard_labels(self, labels, batch):
        """The kernels index `labels` as [N, U-1]: a narrower tensor must never reach them.  That case is a
        length mismatch by the reference's rules, so it is settled on the spot (host round trip) instead of
        after the launch."""
        if labels.shape[0] != batch or labels.shape[1] < self.U - 1:
            self.finish()
            raise ValueError("labels must hold U-1 entries per utterance")

    def finish(self):
        if self.event is not None:
            self.event.synchronize()
            self.event = None
            self.values = self.values.tolist()
        if self.T != self.values[0]:
            raise ValueError("Input length mismatch")
        if self.U != self.values[1] + 1:
            raise ValueError("Output length mismatch")


def certify_inputs(log_probs, labels, lengths, label_lengths, defer=False):
    """defer=False: the reference's behaviour, everything checked before returning (one host round trip).
    defer=True: returns a LengthCheck whose finish() the caller runs AFTER queueing its kernels."""
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
    check = LengthCheck(lengths, label_lengths, log_probs.shape[1], log_probs.shape[2])
    if defer:
        return check
    check.finish()
    return None

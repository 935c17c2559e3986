"""warprnnt_pytorch — RNN-T loss operator for PyTorch on the B200-native libwarprnnt.

Same public surface as the reference package (pytorch_binding/warprnnt_pytorch/__init__.py):
``RNNTLoss(blank=0, reduction='mean')``, ``rnnt_loss(acts, labels, act_lens, label_lens,
blank=0, reduction='mean')`` and the ``warp_rnnt`` extension functions, with the reference's
input rules and error types (certify_inputs, :115-140).  Differences, all on the fast side:
the call never synchronises with the host except for the reference's own length check, costs
stay on the device, and the gradient is produced in autograd's backward with grad_output and the
'mean' factor folded into the kernel - no zeros_like / mul_ passes over the [N,T,U,V] tensor.
CPU tensors are rejected: there is no host path.
"""
import torch
from torch.autograd import Function
from torch.nn import Module

from . import warp_rnnt
from .warp_rnnt import cpu_rnnt, gpu_rnnt, gpu_rnnt_async, gpu_rnnt_backward, gpu_rnnt_forward  # noqa: F401

__all__ = ['rnnt_loss', 'RNNTLoss']


class _RNNT(Function):
    """One training step costs the algorithmic 12 B per logit: forward() reads the logits once
    (statistics + alpha/beta lattices, 4 B), backward() reads them again and writes the gradient
    with the upstream gradient and the 'mean' factor already applied (8 B).  The reference makes
    four more full-tensor passes around its C call (zeros_like, memset, grads /= N, grads.mul_)."""

    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, blank, reduction):
        """
        acts: (batch x seqLength x labelLength x outputDim) raw joint-network logits
        labels: (batch x maxLabelLength) int32 targets, zero padded
        act_lens / label_lens: (batch) int32
        """
        length_check = certify_inputs(acts, labels, act_lens, label_lens, defer=True)
        if not acts.is_cuda:
            raise RuntimeError("warprnnt_pytorch (B200 build) runs on CUDA tensors only; "
                               "there is no CPU fallback")
        warp_rnnt.require_same_device(acts, labels=labels, act_lens=act_lens, label_lens=label_lens)
        if reduction not in ('none', 'sum', 'mean'):
            raise ValueError("reduction must be 'none', 'sum' or 'mean'")
        minibatch_size = acts.size(0)
        length_check.guard_labels(labels, minibatch_size)
        need_grad = acts.requires_grad
        # bf16 / fp16 logits: arithmetic, lattice and costs are fp32 (6 B per logit instead of 12)
        costs = torch.empty(minibatch_size, dtype=warp_rnnt.costs_dtype(acts), device=acts.device)
        ws = warp_rnnt.gpu_rnnt_forward(acts, labels, act_lens, label_lens, costs, blank,
                                        prepare_backward=need_grad)
        length_check.finish()   # T == max(act_lens), U == max(label_lens) + 1: waited for with the kernels queued
        if need_grad:
            ctx.save_for_backward(acts, labels, act_lens, label_lens)
            ctx.workspace = ws
            ctx.blank = blank
            # reference :38-40 divides costs and grads by N for 'mean'
            ctx.scale = 1.0 / minibatch_size if reduction == 'mean' else 1.0
        if reduction in ('sum', 'mean'):
            costs = costs.sum().unsqueeze_(-1)
            if reduction == 'mean':
                costs /= minibatch_size
        return costs

    @staticmethod
    def backward(ctx, grad_output):
        # reference :47-50: grads.mul_(grad_output.view(-1,1,1,1)); here the factor rides in the kernel
        acts, labels, act_lens, label_lens = ctx.saved_tensors
        n = acts.size(0)
        g = grad_output.reshape(-1).to(device=acts.device, dtype=warp_rnnt.costs_dtype(acts))
        g = g.expand(n).contiguous() if g.numel() == 1 else g.contiguous()
        grads = torch.empty_like(acts)   # the kernel defines every element (zeros on padding)
        warp_rnnt.gpu_rnnt_backward(acts, labels, act_lens, label_lens, grads, g, ctx.blank,
                                    ctx.scale, ctx.workspace)
        return grads, None, None, None, None, None


def rnnt_loss(acts, labels, act_lens, label_lens, blank=0, reduction='mean'):
    """RNN Transducer loss (reference :53-70).

    reduction: 'none' | 'sum' | 'mean'; 'mean' divides the summed loss by the batch size (what
    the reference computes, :36-40).
    """
    return _RNNT.apply(acts, labels, act_lens, label_lens, blank, reduction)


class RNNTLoss(Module):
    """Module form (reference :73-100): RNNTLoss(blank=0, reduction='mean')."""

    def __init__(self, blank=0, reduction='mean'):
        super(RNNTLoss, self).__init__()
        self.blank = blank
        self.reduction = reduction
        self.loss = _RNNT.apply

    def forward(self, acts, labels, act_lens, label_lens):
        return self.loss(acts, labels, act_lens, label_lens, self.blank, self.reduction)


from ._checks import certify_inputs, check_contiguous, check_dim, check_type  # noqa: E402,F401

"""`modules.msg_agg.{LastAggregator, MeanAggregator}` (reference
modules/msg_agg.py:15-26) on csrc/agg.cu.

LastAggregator: deterministic first-wins argmax by timestamp + row gather
(the reference's torch_scatter CPU rule; its CUDA rule is nondeterministic).
MeanAggregator: segmented mean, empty segments 0."""
import torch
from torch import Tensor

from tgn_b200 import ops


class _LastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, msg, index, t, dim_size):
        out, argmax = ops.agg_last(msg, index, t, dim_size)
        ctx.save_for_backward(argmax)
        ctx.shape = msg.shape
        return out

    @staticmethod
    def backward(ctx, g):
        (argmax,) = ctx.saved_tensors
        d_msg = g.new_zeros(ctx.shape)
        mask = argmax < ctx.shape[0]
        d_msg[argmax[mask]] = g[mask]
        return d_msg, None, None, None


class _MeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, msg, index, dim_size):
        ctx.save_for_backward(index)
        ctx.dim_size = dim_size
        return ops.agg_mean(msg, index, dim_size)

    @staticmethod
    def backward(ctx, g):
        (index,) = ctx.saved_tensors
        count = torch.bincount(index, minlength=ctx.dim_size).clamp(min=1).to(g.dtype)
        return (g / count.unsqueeze(-1))[index], None, None


class LastAggregator(torch.nn.Module):
    def forward(self, msg: Tensor, index: Tensor, t: Tensor, dim_size: int):
        return _LastFn.apply(msg, index, t, dim_size)


class MeanAggregator(torch.nn.Module):
    def forward(self, msg: Tensor, index: Tensor, t: Tensor, dim_size: int):
        return _MeanFn.apply(msg, index, dim_size)

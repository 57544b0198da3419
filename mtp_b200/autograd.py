"""torch.autograd bridge: one Function for the whole backbone (forward saves activations in engine-owned buffers,
backward runs the hand-written backward kernels and returns fp32 parameter gradients in nn.Module.parameters() order)."""
import torch

from . import engine


class BackboneFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, keep, *params):
        outs, saved = engine._forward_impl(module, x, keep, save=True)
        ctx.module = module
        ctx.saved = saved
        ctx.x = x
        ctx.n_params = len(params)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        from . import engine_bwd
        grads = engine_bwd.backward_impl(ctx.module, ctx.x, ctx.saved, grad_outs)
        ctx.saved = None
        return (None, None, None) + tuple(grads)

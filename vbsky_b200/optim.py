"""Optimiser step of ``SeqData.loop`` on the device (/root/reference/vbsky/fasta.py:388, 421-433): the reference drives
``jax.example_libraries.optimizers.adagrad(step_size)`` (momentum 0.9 on the normalised gradient) with
``nan_to_num``-cleaned gradients; here one fused kernel (``vbsky_adagrad_update``) updates a flat block of variational
parameters in place."""
import ctypes as C

import torch

from . import _lib


class Adagrad:
    """State (sum of squared gradients, momentum buffer) for one flat fp64 CUDA parameter tensor."""

    def __init__(self, x: torch.Tensor, step_size: float = 1.0, momentum: float = 0.9):
        if x.dtype != torch.float64 or not x.is_cuda or not x.is_contiguous():
            raise ValueError("Adagrad needs a contiguous float64 CUDA tensor")
        self.x, self.step_size, self.momentum = x, float(step_size), float(momentum)
        self.g_sq = torch.zeros_like(x)
        self.m = torch.zeros_like(x)

    def update(self, g: torch.Tensor):
        """x <- opt_update(i, nan_to_num(g), state) in place (the step size is constant in the reference)."""
        if g.shape != self.x.shape or g.dtype != torch.float64 or g.device != self.x.device:
            raise ValueError("gradient does not match the parameter block")
        g = g.contiguous()
        dev = self.x.device
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.lib.vbsky_adagrad_update(
                self.x.numel(), C.c_void_p(self.x.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(self.g_sq.data_ptr()),
                C.c_void_p(self.m.data_ptr()), self.step_size, self.momentum, st))

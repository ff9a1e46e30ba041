"""Fused clip_grad_norm_ + SGD(momentum, weight decay) step (reference utils/trainer.py:149-151, train.py:73-78).

FusedSGD is a torch.optim.SGD (same param_groups / state_dict layout, so the reference's checkpoint format holds)
whose gradients and momentum buffers live in two flat fp32 buffers: one memset zeroes all gradients, one kernel pair
(sum of squares, then clip + weight decay + momentum + update) replaces ~700 foreach launches, and the flat gradient
buffer is what the data-parallel path hands to NCCL.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops


class FusedSGD(torch.optim.SGD):
    def __init__(self, params, lr=0.01, momentum=0.9, weight_decay=1e-4, max_norm=1.0):
        super().__init__(params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.max_norm = max_norm
        ps = [p for g in self.param_groups for p in g["params"]]
        if not ps or not ps[0].is_cuda:
            raise RuntimeError("dfcsa.FusedSGD needs CUDA parameters (no CPU fallback)")
        if len(self.param_groups) != 1:
            raise NotImplementedError("dfcsa.FusedSGD supports one param group (the reference uses one)")
        self._params = ps
        dev = ps[0].device
        from .ddp import flat_offsets
        offsets, total = flat_offsets(ps)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_mom = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grads, self._mom = {}, {}
        table = (L.ParamDesc * len(ps))()
        for i, p in enumerate(ps):
            n = p.numel()
            off = offsets[p][0]
            g = self.flat_grad[off:off + n].view(p.shape)
            m = self.flat_mom[off:off + n].view(p.shape)
            self.grads[p] = g
            self._mom[p] = m
            p.grad = g
            table[i].w, table[i].g, table[i].m, table[i].n = p.data_ptr(), g.data_ptr(), m.data_ptr(), n
        raw = bytes(table)
        self._table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self._max_n = max(p.numel() for p in ps)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._first = True

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        for p in self._params:     # keep .grad pointing at the flat views
            if p.grad is None or p.grad.data_ptr() != self.grads[p].data_ptr():
                p.grad = self.grads[p]

    def grad_norm(self):
        """global L2 norm of the gradients of the last step() (device tensor, fp64)."""
        return self._sumsq.sqrt()

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        g = self.param_groups[0]
        for p in self._params:     # gradients produced by autograd elsewhere are folded back into the flat buffer
            if p.grad is not None and p.grad.data_ptr() != self.grads[p].data_ptr():
                self.grads[p].copy_(p.grad)
                p.grad = self.grads[p]
        self._sumsq.zero_()
        n = len(self._params)
        ops.grad_sumsq(self._table, n, self._max_n, self._sumsq)
        ops.sgd_step(self._table, n, self._max_n, self._sumsq, grad_scale, self.max_norm if self.max_norm else 0.0,
                     g["lr"], g["momentum"], g["weight_decay"], self._first)
        if self._first:
            for p in self._params:
                self.state[p]["momentum_buffer"] = self._mom[p]
            self._first = False
        from . import engine
        engine.WEIGHTS_EPOCH += 1       # the kernel rewrote the parameters through raw pointers: folded inference operands are stale
        return None

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        any_buf = False
        for p in self._params:
            buf = self.state.get(p, {}).get("momentum_buffer")
            if buf is not None:
                self._mom[p].copy_(buf)
                self.state[p]["momentum_buffer"] = self._mom[p]
                any_buf = True
        self._first = not any_buf

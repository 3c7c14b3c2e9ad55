"""FusedAdam -- torch.optim.Adam for a FastEGNN model as ONE kernel launch per step.

Drop-in for the optimizer the reference builds in its mains
(`optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)`, main_nbody.py:137) and steps in
utils/train.py:168-170.  torch's Adam walks 119-135 small tensors with ~19 multi-tensor launches per step; here the
parameters are re-pointed into one flat buffer (16-byte aligned slots, in `parameters()` order -- the same layout the
model's backward uses for its flat gradient buffer), so the step is `fegnn_adam_step` over flat arrays.
Arithmetic and skipping rules are torch.optim.Adam's (amsgrad=False, maximize=False): parameters whose `.grad` is
None are left untouched, state and all.  The step counter is a device scalar, so the step can be CUDA-graph captured.
"""
from __future__ import annotations

from typing import Iterable

import torch

from . import _lib as L

lib = L.lib


def _al4(n: int) -> int:
    return (n + 3) & ~3


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam takes a single parameter group (the reference builds one)")
        ps = self.param_groups[0]["params"]
        if not ps:
            raise ValueError("FusedAdam got an empty parameter list")
        dev = ps[0].device
        for p in ps:
            if p.device != dev or p.dtype != torch.float32 or not p.is_cuda:
                raise L.FegnnError("FusedAdam needs float32 CUDA parameters on one device (no CPU fallback)")
        self._offs, tot = [], 0
        for p in ps:
            self._offs.append(tot)
            tot += _al4(p.numel())
        self._n = tot
        self._pflat = torch.zeros(tot, device=dev)
        self._m = torch.zeros(tot, device=dev)
        self._v = torch.zeros(tot, device=dev)
        self._gflat = None                       # only used when the gradients are not one flat buffer already
        self._step = torch.zeros(1, device=dev)
        self._live = torch.zeros(tot, device=dev, dtype=torch.uint8)
        self._live_sig = None
        with torch.no_grad():
            for p, o in zip(ps, self._offs):
                view = self._pflat[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                    # the module keeps its Parameter objects; storage is now the flat buffer

    def flat_grad(self):
        """The ONE buffer behind every parameter's .grad when the gradients are views of a flat buffer laid out like the
        parameters (what the stack's backward hands to autograd), else None.  A data-parallel caller all-reduces this tensor
        in place -- one collective, no flatten / unflatten copies around it (they were 3 of the step's kernels)."""
        ps = [p for g in self.param_groups for p in g["params"]]
        if self._flat_grad_ptr(ps) is None:
            return None
        base = None
        for p in ps:
            g = p.grad
            if g is None:
                continue
            b = g._base if g._base is not None else g
            if base is None:
                base = b
            elif b.data_ptr() != base.data_ptr():
                return None
        return base if base is not None and base.is_contiguous() and base.numel() >= self._n else None

    def _flat_grad_ptr(self, ps):
        """Device pointer of a flat gradient buffer laid out like the parameters, or None."""
        base = None
        for p, o in zip(ps, self._offs):
            g = p.grad
            if g is None:
                continue
            if g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
                return None
            b = g.data_ptr() - 4 * o
            if base is None:
                base = b
            elif b != base:
                return None
        return base if base is not None and base % 16 == 0 else None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        ps = group["params"]
        self._readopt(ps)
        sig = tuple(p.grad is not None for p in ps)
        if not any(sig):
            return loss
        if sig != self._live_sig:                # which parameters are skipped (torch: grad is None)
            if self._live_sig is not None:
                # torch keeps a step counter per parameter; one shared device counter is only the same arithmetic while
                # the set of parameters that receive gradients does not change (true for FastEGNN: the last layer's
                # node_mlp* never do)
                raise L.FegnnError("FusedAdam: the set of parameters with gradients changed between steps; "
                                   "use torch.optim.Adam for such a model")
            live = torch.zeros(self._n, dtype=torch.uint8)
            for p, o, on in zip(ps, self._offs, sig):
                if on:
                    live[o:o + p.numel()] = 1
                    self.state[p] = dict(step=self._step, exp_avg=self._m[o:o + p.numel()].view_as(p),
                                         exp_avg_sq=self._v[o:o + p.numel()].view_as(p))
            self._live.copy_(live)
            self._live_sig = sig
        gptr = self._flat_grad_ptr(ps)
        if gptr is None:                         # foreign gradients: gather them into our own flat buffer
            if self._gflat is None:
                self._gflat = torch.zeros(self._n, device=self._pflat.device)
            dst = [self._gflat[o:o + p.numel()].view_as(p) for p, o, on in zip(ps, self._offs, sig) if on]
            torch._foreach_copy_(dst, [p.grad for p, on in zip(ps, sig) if on])
            gptr = self._gflat.data_ptr()
        b1, b2 = group["betas"]
        st = torch.cuda.current_stream(self._pflat.device).cuda_stream
        with torch.cuda.device(self._pflat.device):          # the C ABI launches on the runtime's current device
            L.check(lib.fegnn_adam_step(self._n, L.ptr(self._pflat), gptr, L.ptr(self._m), L.ptr(self._v),
                                        L.ptr(self._live), L.ptr(self._step), float(group["lr"]), float(b1), float(b2),
                                        float(group["eps"]), float(group["weight_decay"]), st), "fegnn_adam_step")
        return loss

    # -- checkpoint / resume: the kernel reads the flat buffers, so loaded state is copied INTO them (a plain
    #    Optimizer.load_state_dict would leave fresh tensors in self.state that the kernel never sees)
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        ps = self.param_groups[0]["params"]
        steps = set()
        with torch.no_grad():
            for p, o in zip(ps, self._offs):
                st_p = self.state.get(p)
                if not st_p:
                    continue
                m_view = self._m[o:o + p.numel()].view_as(p)
                v_view = self._v[o:o + p.numel()].view_as(p)
                m_view.copy_(st_p["exp_avg"])
                v_view.copy_(st_p["exp_avg_sq"])
                steps.add(float(st_p["step"]))
                self.state[p] = dict(step=self._step, exp_avg=m_view, exp_avg_sq=v_view)
            if len(steps) > 1:
                raise L.FegnnError("FusedAdam.load_state_dict: parameters carry different step counts; one shared "
                                   "device counter cannot represent that")
            if steps:
                self._step.fill_(steps.pop())

    def _readopt(self, ps):
        """model.to() / .float() / _apply after construction give the Parameters fresh storage; point them back into the
        flat buffer (keeping their current values) so the kernel and the model keep seeing the same memory."""
        base = self._pflat.data_ptr()
        with torch.no_grad():
            for p, o in zip(ps, self._offs):
                if p.data_ptr() != base + 4 * o:
                    if p.device != self._pflat.device or p.dtype != torch.float32:
                        raise L.FegnnError("FusedAdam: a parameter left the optimizer's device / dtype")
                    view = self._pflat[o:o + p.numel()].view_as(p)
                    view.copy_(p)
                    p.data = view

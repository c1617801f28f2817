"""Optimiser step of the reference's training loop on the library's multi-tensor kernels (csrc/optim.cu).

Reference (train.py:320-325, utils_train.py:167-181):

    opt = Adam(params, lr=args.learning_rate, weight_decay=args.weight_decay)      # or AdamW(betas=(0.9, 0.95))
    opt.zero_grad(); loss.backward(); clip_grad_norm_(dalle.parameters(), args.clip_grad_norm); opt.step()

`FusedAdam` / `FusedAdamW` subclass `torch.optim.Optimizer` and keep torch's state layout (`step`, `exp_avg`,
`exp_avg_sq` per parameter), so `opt.state_dict()` checkpoints written by train.py:352,387 round-trip with the stock
optimisers.  `clip_grad_norm_` has the signature and the in-place effect of `torch.nn.utils.clip_grad_norm_` (L2 only)
but never synchronises with the host.  Each call is ONE launch over every parameter tensor of the model.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops

CHUNK = 65536  # elements per block; multiple of 4


class _TensorTable:
    """Device-resident (param, grad, exp_avg, exp_avg_sq, n) records + chunk map for a fixed list of tensors."""

    def __init__(self, params, ms=None, vs=None):
        self.params = list(params)
        assert self.params, "empty parameter list"
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise RuntimeError("mmvid_b200.optim: parameters must be contiguous fp32 tensors on one CUDA device")
        if dev.type != "cuda":
            raise RuntimeError("mmvid_b200.optim needs CUDA parameters (there is no CPU fallback)")
        self.device = dev
        self.ms, self.vs = ms, vs
        ct, ci = [], []
        for ti, p in enumerate(self.params):
            for c in range((p.numel() + CHUNK - 1) // CHUNK):
                ct.append(ti)
                ci.append(c)
        self.n_chunks = len(ct)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
        self.chunk_index = torch.tensor(ci, dtype=torch.int32, device=dev)
        self.partial = torch.empty(self.n_chunks, dtype=torch.float32, device=dev)
        self.total_sq = torch.empty(1, dtype=torch.float32, device=dev)
        self._host = (L.AdamTensor * len(self.params))()
        self._pinned = torch.empty(C.sizeof(self._host), dtype=torch.uint8).pin_memory()
        self.table = torch.empty(C.sizeof(self._host), dtype=torch.uint8, device=dev)
        self._grad_ptrs = None
        self.skipped = [0] * len(self.params)

    def refresh(self):
        """Re-upload the table when any pointer in it changed: .grad (zero_grad(set_to_none=True) reallocates them), a
        replaced p.data, or replaced state tensors."""
        ptrs = tuple(0 if p.grad is None else p.grad.data_ptr() for p in self.params) + tuple(self.skipped) \
            + tuple(p.data_ptr() for p in self.params) \
            + (() if self.ms is None else tuple(t.data_ptr() for t in self.ms) + tuple(t.data_ptr() for t in self.vs))
        if ptrs == self._grad_ptrs:
            return
        for i, p in enumerate(self.params):
            e = self._host[i]
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                raise RuntimeError("mmvid_b200.optim: gradients must be contiguous fp32")
            e.p, e.g, e.n = p.data_ptr(), (None if g is None else g.data_ptr()), p.numel()
            e.m = None if self.ms is None else self.ms[i].data_ptr()
            e.v = None if self.vs is None else self.vs[i].data_ptr()
            e.skipped = self.skipped[i]
        C.memmove(self._pinned.data_ptr(), C.addressof(self._host), C.sizeof(self._host))
        self.table.copy_(self._pinned)  # a few KB, synchronous: the pinned staging buffer is reused
        self._grad_ptrs = ptrs

    def args(self):
        return (ops._ptr(self.table), ops._ptr(self.chunk_tensor), ops._ptr(self.chunk_index), self.n_chunks, CHUNK)


_clip_tables = {}


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """torch.nn.utils.clip_grad_norm_ (train.py:324) for the L2 norm; returns the total norm as a 0-d device tensor."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("only the L2 norm the reference uses")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return torch.zeros(())
    key = tuple(id(p) for p in params)
    tab = _clip_tables.get(key)
    if tab is None:
        _clip_tables.clear()  # one model at a time; do not pin stale parameter lists
        tab = _clip_tables[key] = _TensorTable(params)
    tab.refresh()
    lib = L.load()
    L.check(lib.mmvid_grad_sqnorm(*tab.args(), ops._ptr(tab.partial), ops._ptr(tab.total_sq), ops._stream()), "grad_sqnorm")
    L.check(lib.mmvid_grad_clip(*tab.args(), ops._ptr(tab.total_sq), float(max_norm), ops._stream()), "grad_clip")
    return tab.total_sq.sqrt()[0]


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) semantics (no amsgrad / maximize)."""

    decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._tables = {}

    def _table(self, gi, group):
        ps = [p for p in group["params"] if p.requires_grad]
        key = (gi, tuple(id(p) for p in ps))
        tab = self._tables.get(key)
        if tab is None:
            ms, vs = [], []
            for p in ps:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            tab = self._tables[key] = _TensorTable(ps, ms, vs)
        return tab

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}  # the state tensors were replaced

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            tab = self._table(gi, group)
            # torch keeps one step counter per parameter; a parameter without a gradient does not advance.  The launch
            # carries the largest counter, each record how far its tensor lags behind it.
            for p in tab.params:
                if p.grad is not None:
                    self.state[p]["step"] += 1
            steps = [int(self.state[p]["step"]) for p in tab.params]
            step = max(steps)
            tab.skipped = [step - s for s in steps]
            tab.refresh()
            b1, b2 = group["betas"]
            L.check(lib.mmvid_adam_step(*tab.args(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                        float(group["weight_decay"]), int(self.decoupled), step, ops._stream()), "adam_step")
            # the kernel wrote the parameters through raw pointers: torch's version counters did not move
            L.bump_weights_epoch()
        return loss


class FusedAdamW(FusedAdam):
    """torch.optim.AdamW semantics (decoupled weight decay; utils_train.py:173-179 uses betas=(0.9, 0.95))."""

    decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)


def get_optimizer(optimizer, params, learning_rate, weight_decay):
    """utils_train.py:167-181 with the fused optimisers."""
    if optimizer == "adam":
        return FusedAdam(params, lr=learning_rate, weight_decay=weight_decay)
    if optimizer == "adamw":
        return FusedAdamW(params, lr=learning_rate, betas=(0.9, 0.95), weight_decay=weight_decay)
    raise NotImplementedError(optimizer)

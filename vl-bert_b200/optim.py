"""Fused optimizer step for the path's parameters: the reference's AdamW (common/nlp/bert/optimization.py:107-187) and the
trainer's global-norm gradient clip (common/trainer.py:139-147) as two kernel launches over a descriptor table of every
parameter tensor (vlb_grad_sqnorm + vlb_adamw_step) instead of ~10 elementwise kernels per tensor.

    opt = vlbert_b200.optim.FusedAdamW(param_groups, lr=..., betas=(0.9, 0.999), eps=1e-6, weight_decay=..., max_grad_norm=1.0)
    loss.backward(); opt.step(); opt.zero_grad()

Constructor arguments, param-group keys, `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq` per parameter) are those of
the reference class, so its checkpoints and LR schedulers (`WarmupLinearSchedule`, ... :14-104) work unchanged.
`max_grad_norm` replaces the separate `clip_grad_norm_` call: the clipping coefficient is computed on the device and applied
inside the update (the stored gradients are not modified); `last_total_norm` holds the norm (device scalar) for logging.
No CPU path: parameters must be CUDA fp32 tensors.
"""
import ctypes
import math

import torch
from torch.optim import Optimizer

from . import _lib


class FusedAdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True, max_grad_norm=0.0):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {}".format(betas))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))
        self.max_grad_norm = float(max_grad_norm)
        self._key = None
        self._table = None
        self._hyper_host = None
        self._hyper_dev = None
        self._sq = None
        self._copied = None
        self.last_total_norm = None

    def _plan(self):
        """(group, param) pairs with a gradient, grouped by (betas, eps) since those are kernel-wide constants"""
        plan = {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32):
                    raise RuntimeError("vlbert_b200.FusedAdamW: parameters and gradients must be CUDA fp32 tensors (there is no CPU path)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("vlbert_b200.FusedAdamW: parameters and gradients must be contiguous")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p.data)
                    state["exp_avg_sq"] = torch.zeros_like(p.data)
                plan.setdefault((group["betas"][0], group["betas"][1], group["eps"]), []).append((group, p))
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        plan = self._plan()
        if not plan:
            return loss
        lib = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        flat = [(consts, g, p) for consts, items in plan.items() for g, p in items]
        dev = flat[0][2].device
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for _, _, p in flat)
        if key != self._key:
            arr = (_lib.AdamWTensor * len(flat))()
            for i, (_, _, p) in enumerate(flat):
                s = self.state[p]
                arr[i].param, arr[i].grad = p.data_ptr(), p.grad.data_ptr()
                arr[i].exp_avg, arr[i].exp_avg_sq, arr[i].n = s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr(), p.numel()
            self._table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
            self._hyper_host = torch.empty((len(flat), 4), dtype=torch.float32).pin_memory()
            self._hyper_dev = torch.empty((len(flat), 4), dtype=torch.float32, device=dev)
            self._sq = torch.zeros((1,), dtype=torch.float32, device=dev)
            self._key = key
        vals = []
        for _, group, p in flat:
            s = self.state[p]
            s["step"] += 1
            lr, (b1, b2) = group["lr"], group["betas"]
            step_size = lr
            if group["correct_bias"]:
                step_size = step_size * math.sqrt(1.0 - b2 ** s["step"]) / (1.0 - b1 ** s["step"])
            vals.append((lr, lr * group["weight_decay"], step_size, 0.0))
        if self._copied is not None:
            self._copied.synchronize()          # the previous step's upload must have left the pinned buffer
        self._hyper_host.copy_(torch.tensor(vals, dtype=torch.float32))
        self._hyper_dev.copy_(self._hyper_host, non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record()
        esz = ctypes.sizeof(_lib.AdamWTensor)
        sq_ptr = None
        if self.max_grad_norm > 0:
            _lib.check(lib.vlb_grad_sqnorm(self._table.data_ptr(), len(flat), self._sq.data_ptr(), st))
            sq_ptr = self._sq.data_ptr()
            self.last_total_norm = self._sq.sqrt()
        o = 0
        for consts, items in plan.items():      # one launch per distinct (betas, eps): normally exactly one
            _lib.check(lib.vlb_adamw_step(self._table.data_ptr() + o * esz, self._hyper_dev.data_ptr() + o * 16, len(items),
                                          float(consts[0]), float(consts[1]), float(consts[2]), sq_ptr, self.max_grad_norm, st))
            o += len(items)
        return loss

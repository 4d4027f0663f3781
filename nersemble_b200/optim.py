"""Optimiser for the `fields` parameter group with the hash-table update fused into one CUDA pass.

The reference trains with nerfstudio's AdamOptimizerConfig (torch.optim.Adam, eps = 1e-15) over the 8 dense tcnn grid
gradients (train_nersemble.py: optimizers["fields"]).  For the 403 M-element table that costs, per step, the dense
expansion of the scattered gradient (3.1 ms), torch's multi-pass foreach Adam (6.0 ms) and a fresh fp32 -> fp16 copy
of the table for the next forward (r1d profile, 1 x B200).  `FusedFieldsAdam` is a torch.optim.Adam whose update of
the table parameter is nsb_table_adam_step: rank-1 gradient expansion + Adam + fp16 refresh in one streaming pass;
every other parameter of the group goes through torch.optim.Adam unchanged, and the optimiser state keeps torch's
keys (`step`, `exp_avg`, `exp_avg_sq`), so checkpoints interchange with a plain Adam.

    opt = FusedFieldsAdam(model.get_param_groups()["fields"], lr=5e-3, eps=1e-15)     # finds the table by itself

nerfstudio: `FusedFieldsAdamOptimizerConfig` has AdamOptimizerConfig's fields and `setup(params)` signature.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops


class FusedFieldsAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, **kw):
        if kw.get("amsgrad") or kw.get("maximize"):
            raise NotImplementedError("FusedFieldsAdam: amsgrad / maximize")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)
        # torch.cuda.amp.GradScaler unscales `.grad` tensors only; the parked table gradient is not one.  A trainer that
        # scales the loss sets this to 1 / scaler.get_scale() before step() (Adam is scale-invariant up to eps, so
        # forgetting it changes the update only where |g| ~ eps = 1e-15).
        self.pending_grad_scale = 1.0
        self._ensembles = []
        for group in self.param_groups:
            for p in group["params"]:
                he = getattr(p, "_nsb_hash_ensemble", None)
                he = he() if he is not None else None
                if he is not None:
                    he.defer_table_grad = True     # backward leaves the table gradient in rank-1 form
                    self._ensembles.append((he, p, group))
        if not self._ensembles:
            raise ValueError("FusedFieldsAdam: no HashEnsemble table among the parameters (use torch.optim.Adam)")

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none)
        for he, _, _ in self._ensembles:
            he.pending_table_grad = None

    @torch.no_grad()
    def step(self, closure=None):
        work = []
        for he, p, group in self._ensembles:
            if p.grad is not None and he.pending_table_grad is not None:
                # a dense .grad AND a parked rank-1 gradient (second backward before step()): they may carry different
                # scale factors (all-reduce averaging), so fold the parked one into the dense tensor first
                he.materialize_pending(scale=float(self.pending_grad_scale))
            work.append((he, p, group, p.grad, he.pending_table_grad))
            p.grad = None                   # torch's Adam skips parameters without .grad
        loss = super().step(closure)
        for he, p, group, dense, pending in work:
            if dense is None and pending is None:
                continue
            if not p.is_cuda:
                raise RuntimeError("FusedFieldsAdam needs the table on a CUDA device (there is no CPU fallback)")
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["step"] += 1
            shadow = he.shadow_buffer()
            ops.table_adam_step(p.data, st["exp_avg"], st["exp_avg_sq"], shadow, step=int(st["step"].item()),
                                lr=float(group["lr"]), betas=group["betas"], eps=group["eps"],
                                weight_decay=group["weight_decay"], grad=dense, pending=pending,
                                grad_scale=(1.0 if pending is None else float(pending.get("scale", 1.0)) * float(self.pending_grad_scale)))
            torch.autograd.graph.increment_version(p)
            he.set_native_tables(shadow)
            he.pending_table_grad = None
        return loss


@dataclass
class FusedFieldsAdamOptimizerConfig:
    """Drop-in for nerfstudio.engine.optimizers.AdamOptimizerConfig in the method config's `fields` entry."""
    lr: float = 5e-3
    eps: float = 1e-15
    max_norm: Optional[float] = None
    weight_decay: float = 0.0

    def setup(self, params) -> FusedFieldsAdam:
        return FusedFieldsAdam(params, lr=self.lr, eps=self.eps, weight_decay=self.weight_decay)

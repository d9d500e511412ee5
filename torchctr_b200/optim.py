"""``FusedAdagrad``: ``torch.optim.Adagrad`` whose step over the dense (tower) parameters is ONE kernel launch.

The reference hands ``optimizer.step()`` a handful of small dense tensors besides the tables
(``torchctr/trainer.py:303``); torch's foreach implementation spends five launches on them.  Same arithmetic, same
``state_dict`` (``sum`` / ``step`` per parameter), so it can replace ``torch.optim.Adagrad`` in an existing script; the
embedding tables are expected to go through ``model.bind_optimizer`` (they never get a ``.grad``)."""
from __future__ import annotations

import torch

from . import ops


class FusedAdagrad(torch.optim.Adagrad):
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            fusable = (group.get("lr_decay", 0) == 0 and group.get("weight_decay", 0) == 0 and not group.get("maximize", False)
                       and all(p.is_cuda and p.dtype == torch.float32 and not p.grad.is_sparse and p.is_contiguous()
                               and p.grad.is_contiguous() for p in params))
            if not fusable or not params:
                self._torch_group_step(group)
                continue
            sums = []
            for p in params:
                st = self.state[p]
                st["step"] += 1                      # host tensor, as in torch (only lr_decay reads it)
                sums.append(st["sum"])
            for i in range(0, len(params), 48):
                ops.dense_adagrad(params[i:i + 48], [p.grad for p in params[i:i + 48]], sums[i:i + 48],
                                  float(group["lr"]), float(group["eps"]))
        return loss

    def _torch_group_step(self, group):
        """torch's own implementation for one group (sparse gradients, lr_decay, weight decay ...)."""
        saved = self.param_groups
        try:
            self.param_groups = [group]
            torch.optim.Adagrad.step(self)
        finally:
            self.param_groups = saved

"""Oracle (test infrastructure): optimizer updates restricted to touched table rows.

The reference Trainer calls ``optimizer.step()`` on whatever ``torch.optim``
object the user passes (``/root/reference/torchctr/trainer.py:303``).  These
functions restate, on the unique rows of one step, the torch optimizers whose
result a touched-rows-only update can equal exactly (hard
part 2); ``tests/test_oracle_golden.py`` checks each against the real
``torch.optim`` class:

* ``adagrad_rows``      == ``torch.optim.Adagrad(lr, eps, lr_decay=0, weight_decay=0)``
  on the dense gradient (untouched rows have g = 0 and do not move).
* ``sparse_adam_rows``  == ``torch.optim.SparseAdam`` (lazy Adam: moments of
  untouched rows do not decay).
* ``sgd_rows``          == ``torch.optim.SGD(lr)`` without momentum.
* ``rowwise_adagrad_rows`` -- one accumulator per row, ``state += mean(g*g)``;
  FBGEMM-style, no torch class: parity unpinned by the reference.

All update ``weight`` / state tensors in place and return nothing.
"""
from __future__ import annotations

import math

import torch


def adagrad_rows(weight, state_sum, rows, grads, lr: float, eps: float = 1e-10):
    acc = state_sum[rows] + grads * grads
    state_sum[rows] = acc
    weight[rows] = weight[rows] - lr * (grads / (acc.sqrt() + eps))


def rowwise_adagrad_rows(weight, state_row, rows, grads, lr: float, eps: float = 1e-10):
    acc = state_row[rows] + (grads * grads).mean(dim=1)
    state_row[rows] = acc
    weight[rows] = weight[rows] - lr * (grads / (acc.sqrt() + eps).unsqueeze(1))


def sparse_adam_rows(weight, exp_avg, exp_avg_sq, rows, grads, lr: float, step: int,
                     beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    m_old, v_old = exp_avg[rows], exp_avg_sq[rows]
    m_new = m_old + (grads - m_old) * (1.0 - beta1)
    v_new = v_old + (grads * grads - v_old) * (1.0 - beta2)
    exp_avg[rows] = m_new
    exp_avg_sq[rows] = v_new
    step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    weight[rows] = weight[rows] - step_size * (m_new / (v_new.sqrt() + eps))


def sgd_rows(weight, rows, grads, lr: float):
    weight[rows] = weight[rows] - lr * grads

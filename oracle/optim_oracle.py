"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the optimizer step that follows the hot path.

Follows (reference file:line):
  AdamW.step                      common/nlp/bert/optimization.py:129-187
  torch.nn.utils.clip_grad_norm_  as called by common/trainer.py:139-147 (norm_type 2)

Pinned against the live reference class in tests/test_oracle_vs_reference.py and by tests/golden/adamw.npz
(oracle/make_golden.py:golden_adamw).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.
"""
import math

import numpy as np


def clip_coef(grads, max_norm):
    """-> (coefficient applied to every gradient, total_norm).  clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1."""
    total = math.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
    if max_norm is None or max_norm <= 0:
        return 1.0, total
    return min(1.0, max_norm / (total + 1e-6)), total


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0, correct_bias=True):
    """One parameter tensor, fp32 arithmetic in the reference's order; arrays are updated in place; `step` is the 1-based count."""
    f = np.float32
    m *= f(beta1)
    m += f(1.0 - beta1) * g
    v *= f(beta2)
    v += f(1.0 - beta2) * g * g
    denom = np.sqrt(v) + f(eps)
    step_size = lr
    if correct_bias:
        step_size = step_size * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p += f(-step_size) * (m / denom)
    if weight_decay > 0.0:
        p += f(-lr * weight_decay) * p
    return p

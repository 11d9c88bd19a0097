"""TEST INFRASTRUCTURE ONLY -- the optimiser the reference drives in SeqData.loop (fasta.py:388, 421-433):
``jax.example_libraries.optimizers.adagrad(step_size, momentum=0.9)`` on ``nan_to_num`` gradients.  JAX is absent from
this image; the update rule is restated from its published source (parity unpinned by reference outputs)."""
import numpy as np


def adagrad_init(x0):
    return np.array(x0, dtype=np.float64), np.zeros_like(x0, dtype=np.float64), np.zeros_like(x0, dtype=np.float64)


def adagrad_update(g, state, step_size, momentum=0.9):
    x, g_sq, m = state
    g = np.nan_to_num(np.asarray(g, dtype=np.float64))  # fasta.py:421,427
    with np.errstate(divide="ignore", over="ignore"):
        g_sq = g_sq + np.square(g)
        g_sq_inv_sqrt = np.where(g_sq > 0, 1.0 / np.sqrt(g_sq), 0.0)
    m = (1.0 - momentum) * (g * g_sq_inv_sqrt) + momentum * m
    x = x - step_size * m
    return x, g_sq, m

"""TEST INFRASTRUCTURE (oracle): the dropout mask of the CUDA path, restated on the CPU.

The reference draws its dropout masks from torch's Philox stream (nn.Dropout inside the differentiated path,
models/modules.py:51-52, 236-237); no other implementation can reproduce that stream, and nothing in the result depends
on WHICH Bernoulli(p) mask is drawn -- only on the same mask multiplying the value and every derivative sweep of a point
(they are autograd derivatives of one graph).  The CUDA path draws its masks from a counter-based hash of
(step seed, layer salt, row, column) (porous-cfd_b200/csrc/common.cuh: dropout_seed_hash / dropout_row_hash /
dropout_from_row).  To compare a TRAINING-mode step with the oracle, this module recomputes exactly those masks with
numpy integer arithmetic and the oracle's forward multiplies by them where the reference applies nn.Dropout; autograd
then carries the mask through every sweep, as in the reference.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.
"""
from __future__ import annotations

import math

import numpy as np
import torch

M32 = np.uint64(0xFFFFFFFF)


def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7feb352d)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846ca68b)) & M32
    x ^= x >> np.uint64(16)
    return x


def seed_hash(seed: int, salt: int) -> np.ndarray:
    seed &= 0xFFFFFFFFFFFFFFFF
    lo, hi = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    inner = _mix32(np.array((int(hi) + salt) & 0xFFFFFFFF, dtype=np.uint64))
    return _mix32(np.array(int(lo) ^ int(inner), dtype=np.uint64))


def keep_scale(seed: int, salt: int, rows: np.ndarray, n_cols: int, p: float) -> np.ndarray:
    """(len(rows), n_cols) float32: 1/(1-p) where element (row, col) is kept, 0 where it is dropped."""
    hseed = seed_hash(seed, salt)
    rows = rows.astype(np.uint64)
    hrow = _mix32(hseed ^ (rows & M32))                                     # dropout_row_hash
    cols = np.arange(n_cols, dtype=np.uint64)
    x = (hrow[:, None] + np.uint64(0x9e3779b9) * cols[None, :] + (rows >> np.uint64(32))[:, None]) & M32
    h = _mix32(x)
    thr = np.uint64(math.ceil(np.float32(p) * np.float32(16777216.0)))     # the kernel's ceilf(p * 2^24) in fp32
    keep = (h >> np.uint64(8)) >= thr
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(keep, inv, np.float32(0.0)).astype(np.float32)


class CounterDropout:
    """Callable handed to pinn_oracle.forward(..., dropout_fn=...): multiplies the activations of all N = NI + NB points
    (internal rows first) by the masks the CUDA step uses.  `seed` = the device step seed AFTER the step advanced it
    (executor.ctx.seed_dev), `site` = index in the model's point chain of the layer that CONSUMES the dropped activations
    (the kernels apply activation / dropout while loading a layer's input); the internal and the boundary chain use the
    salts 100 + site and 200 + site and number their rows per chain (engine.PinnExecutor.step)."""

    def __init__(self, seed: int, n_internal: int, n_boundary: int):
        self.seed, self.ni, self.nb = int(seed), n_internal, n_boundary

    def __call__(self, x: torch.Tensor, p: float, site: int) -> torch.Tensor:
        b, n, c = x.shape
        assert n == self.ni + self.nb
        g = np.arange(b, dtype=np.uint64)
        ri = (g[:, None] * np.uint64(self.ni) + np.arange(self.ni, dtype=np.uint64)[None, :]).reshape(-1)
        rb = (g[:, None] * np.uint64(self.nb) + np.arange(self.nb, dtype=np.uint64)[None, :]).reshape(-1)
        mi = keep_scale(self.seed, 100 + site, ri, c, p).reshape(b, self.ni, c)
        mb = keep_scale(self.seed, 200 + site, rb, c, p).reshape(b, self.nb, c)
        mask = torch.from_numpy(np.concatenate([mi, mb], axis=1)).to(x.dtype)
        return x * mask

"""Feature scalers and batch collation (reference dataset/foam_dataset.py:17-90).  The OpenFOAM
case parser and the stratified sampler of the reference are out of scope (SURVEY.md section 2
rows 8-9): batches come from porous_cfd_b200.synthetic or from the reference's own FoamDataset."""
from __future__ import annotations

import numpy as np
import torch

from .foam_data import FoamData


def _as_tensor(v, *args, **kwargs):
    t = v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))
    return t.to(*args, **kwargs)


class StandardScaler:
    """z-score scaling; works on numpy arrays and, after `.to(...)`, on tensors."""

    def __init__(self, std, mean):
        self.std, self.mean = std, mean

    def transform(self, data):
        return (data - self.mean) / self.std

    def inverse_transform(self, data):
        return self.std * data + self.mean

    def __getitem__(self, item):
        return StandardScaler(self.std[item], self.mean[item])

    def to(self, *args, **kwargs) -> 'StandardScaler':
        self.std, self.mean = _as_tensor(self.std, *args, **kwargs), _as_tensor(self.mean, *args, **kwargs)
        return self


class Normalizer:
    """min-max scaling to [0, 1]."""

    def __init__(self, min, max):
        self.min, self.max = min, max
        self.range = max - min

    def transform(self, data):
        return (data - self.min) / self.range

    def inverse_transform(self, data):
        return self.min + self.range * data

    def __getitem__(self, item):
        return Normalizer(self.min[item], self.max[item])

    def to(self, *args, **kwargs) -> 'Normalizer':
        self.min, self.max = _as_tensor(self.min, *args, **kwargs), _as_tensor(self.max, *args, **kwargs)
        self.range = _as_tensor(self.range, *args, **kwargs)
        return self


def collate_fn(samples: list) -> FoamData:
    """Stack per-geometry FoamData into one batch (data (B,N,F), every sub-domain (B,n))."""
    first = samples[0]
    data = torch.stack([s.data for s in samples])
    domain = {name: torch.stack([s.domain[name] for s in samples]) for name in first.domain}
    return FoamData(data, first.labels, domain)

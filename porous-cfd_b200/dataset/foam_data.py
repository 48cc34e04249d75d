"""FoamData: a labelled, sub-domain aware view of one (batched) point-cloud tensor.

Same contract as the reference container (dataset/foam_data.py:5-90): `fd['U']` selects the
columns of a label (single label = its position among the dict keys, multi label = concatenation
of its sub-labels), `fd['internal']` selects the rows of a sub-domain and returns a new FoamData
whose only domain entry is that sub-domain re-indexed from zero.

Label / sub-domain selection on this class is host-side convenience (plain tensor indexing); the
training step does not go through it -- the kernels receive `data`, the column indices and the
row-id tensors directly (pcfd_gather_cols / pcfd_residual_loss fold the gathers into addressing).
"""
from __future__ import annotations

import torch
from torch import Tensor


class FoamData:
    def __init__(self, data: Tensor, labels: dict, domain: dict):
        self.data = data
        self.labels = labels
        self.domain = domain
        # optional: the batch's set-abstraction geometry (FPS centroids, edge slots, centroid positions per level) when a
        # DeviceFoamDataset with a geometry cache built it; the training step then skips FPS / ball query
        self.geometry = None

    # ---- column bookkeeping ---------------------------------------------------------------
    def columns(self, label: str) -> list[int]:
        """Column indices a label refers to."""
        sub = self.labels[label]
        names = sub if sub else [label]
        order = list(self.labels.keys())
        return [order.index(n) for n in names]

    def __contains__(self, item) -> bool:
        return item in self.labels or item in self.domain

    def __getitem__(self, item):
        if item in self.labels:
            cols = self.columns(item)
            if len(cols) == 1:
                return self.data[..., cols[0]:cols[0] + 1]
            if cols == list(range(cols[0], cols[0] + len(cols))):
                return self.data[..., cols[0]:cols[0] + len(cols)]
            return self.data[..., cols]
        if item in self.domain:
            ids = self.domain[item]
            if self.data.dim() > 2:
                picked = torch.gather(self.data, 1, ids.unsqueeze(-1).expand(-1, -1, self.data.shape[-1]))
            else:
                picked = self.data[ids]
            return FoamData(picked, self.labels, {item: torch.arange(0, len(ids))})
        raise KeyError(f'{item} not found in labels or subdomains. Available labels are '
                       f'{list(self.labels.keys())}. Available subdomains are {list(self.domain.keys())}.')

    # ---- tensor plumbing ------------------------------------------------------------------
    def _map(self, fn) -> 'FoamData':
        return FoamData(fn(self.data), self.labels, {k: fn(v) for k, v in self.domain.items()})

    def squeeze(self) -> 'FoamData':
        return self._map(lambda t: t.squeeze())

    def to(self, *args, **kwargs) -> 'FoamData':
        data = self.data.to(*args, **kwargs)
        # row ids stay int64 whatever dtype the data is cast to
        kw = {k: v for k, v in kwargs.items() if k != 'dtype'}
        ar = [a for a in args if not isinstance(a, torch.dtype)]
        return FoamData(data, self.labels, {k: v.to(*ar, **kw) for k, v in self.domain.items()})

    def detach(self) -> 'FoamData':
        return self._map(lambda t: t.detach())

    def pin_memory(self) -> 'FoamData':
        return self._map(lambda t: t.pin_memory())

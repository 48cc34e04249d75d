"""A dataset resident in HBM: per-case features and batch collation on the device.

The reference keeps every sampled case in host memory (`FoamDataset.data`, dataset/foam_dataset.py:163-165), adds the
signed-distance and boundary-id features with scipy / scikit-learn while loading (`add_features`, :397-404) and builds a
batch by stacking host tensors (`collate_fn`, :83-90) that Lightning then copies to the GPU.  A B200 holds such a
dataset whole (config 2: 5 000 geometries x 3 200 points x 11 floats = 0.7 GB of 180 GB), so here the cases are uploaded
once and

  * `add_sdf` / `add_boundary_id` compute the two features for ALL geometries in one launch each (csrc/ingest.cu),
  * `batch(ids)` is a device gather of the chosen geometries (data and every sub-domain's row ids): the training step
    gets its FoamData without a host->device copy.

The OpenFOAM parser, the normalisation statistics and the stratified sampler stay with the reference (out of scope,
DESIGN.md section 7): the constructor takes what `load_case` produced.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import Tensor

from .. import ops
from .foam_data import FoamData


class DeviceFoamDataset:
    def __init__(self, data: Tensor, labels: dict, domain: dict, device='cuda'):
        """data (G, N, F) float32, domain {name: (G, n) int64 row ids}; internal rows first (load_case order)."""
        if data.dim() != 3:
            raise ValueError('data must be (geometries, points, features)')
        self.data = data.to(device=device, dtype=torch.float32).contiguous()
        self.labels = labels
        self.domain = {k: v.to(device=device, dtype=torch.int64).contiguous() for k, v in domain.items()}

    @classmethod
    def from_samples(cls, samples: Sequence[FoamData], device='cuda') -> 'DeviceFoamDataset':
        """From the per-case FoamData list a FoamDataset holds (`FoamDataset.data`)."""
        first = samples[0]
        data = torch.stack([s.data for s in samples])
        domain = {k: torch.stack([s.domain[k] for s in samples]) for k in first.domain}
        return cls(data, first.labels, domain, device)

    def __len__(self) -> int:
        return self.data.shape[0]

    def _columns(self, label: str) -> list:
        return FoamData(self.data, self.labels, {}).columns(label)

    def _n_internal(self) -> int:
        return int(self.domain['internal'].shape[1])

    # ---- features (reference: FoamDataset.add_features) -------------------------------------------------------
    def add_sdf(self, coord_scale: Optional[Tensor] = None) -> None:
        """Fill label 'sdf' for every geometry (FoamDataset.add_sdf).  coord_scale = the `range` (Normalizer) or `std`
        (StandardScaler) of the coordinate scaler when coordinates are stored normalised."""
        pos = self._columns('C')
        if pos != list(range(pos[0], pos[0] + len(pos))):
            raise ValueError('coordinate columns must be contiguous')
        region = self._columns('cellToRegion')[0] if 'cellToRegion' in self.labels else -1
        if coord_scale is not None:
            coord_scale = torch.as_tensor(coord_scale, dtype=torch.float32).to(self.data.device)
        ops.sdf_feature(self.data, self._n_internal(), pos[0], len(pos), region, self._columns('sdf')[0], coord_scale)

    def add_boundary_id(self, boundary_class: Tensor) -> None:
        """Fill the 'boundaryId' columns (FoamDataset.add_boundary_id).  boundary_class (G, N - n_internal): position
        of each boundary row's patch name in the sorted patch names (the category order of OneHotEncoder)."""
        cols = self._columns('boundaryId')
        if cols != list(range(cols[0], cols[0] + len(cols))):
            raise ValueError('boundaryId columns must be contiguous')
        cls = boundary_class.to(device=self.data.device, dtype=torch.int32).contiguous()
        ops.boundary_one_hot(self.data, self._n_internal(), cls, len(cols), cols[0])

    # ---- batches (reference: collate_fn) -----------------------------------------------------------------------
    def _gather(self, ids: Tensor) -> FoamData:
        names = list(self.domain)
        out = ops.gather_blocks_multi([self.data] + [self.domain[k] for k in names], ids)     # one launch
        return FoamData(out[0], self.labels, dict(zip(names, out[1:])))

    def batch(self, geometry_ids) -> FoamData:
        """collate_fn of the chosen geometries.  Ids given on the host are range-checked; a CUDA id tensor is used as
        it is (checking it would cost a device synchronisation per batch)."""
        if torch.is_tensor(geometry_ids) and geometry_ids.is_cuda:
            return self._gather(geometry_ids.to(torch.int64))
        ids = torch.as_tensor(geometry_ids, dtype=torch.int64)
        if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= len(self)):
            raise IndexError('geometry id out of range')
        return self._gather(ids.to(self.data.device))

    def batches(self, batch_size: int, shuffle: bool = True, generator: Optional[torch.Generator] = None,
                drop_last: bool = False):
        """One epoch of batches (the DataLoader + collate_fn of the reference's training scripts)."""
        g = len(self)
        order = torch.randperm(g, generator=generator) if shuffle else torch.arange(g)
        order = order.to(self.data.device)
        for lo in range(0, g, batch_size):
            ids = order[lo:lo + batch_size]
            if drop_last and ids.numel() < batch_size:
                return
            yield self._gather(ids)

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
        self.geo_cache = None

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

    # ---- per-geometry cache of the set-abstraction geometry (second half) ----------------------
    def build_geometry_cache(self, model, chunk: int = 64) -> None:
        """FPS centroids and ball-query neighbourhoods of every resident geometry, for `model`'s set-abstraction stack.

        The reference samples its point clouds once (`FoamDataset.__init__`, dataset/foam_dataset.py:159-161), so the
        encoder's FPS / radius results of a geometry are the same in every epoch -- PROVIDED the FPS start is fixed.
        `torch_cluster.fps` defaults to a random start (which the reference does not override); this build uses the first
        point (oracle/pyg_restate.py), and only under that documented deviation is the cache exact.  Opt-in: batches made
        after this call carry `FoamData.geometry`, which the training step uses instead of running FPS / ball query.
        Indices are stored local to their geometry; the edge slots depend on a geometry's position in the batch (PyG's
        bipartite self-loop rule) and are rebuilt per batch by one launch per level (pcfd_sa_cached_geometry)."""
        from ..engine import sa_geometry_levels
        ex = model.executor
        if not ex.uses_geometry():
            raise ValueError('this model has no set-abstraction encoder: nothing to cache')
        stack = ex.plan['sa_stack']
        g_total = len(self)
        levels = None
        for lo in range(0, g_total, chunk):
            sl = slice(lo, min(g_total, lo + chunk))
            dom = {k: v[sl] for k, v in self.domain.items()}
            pos = ex.geometry_positions(self.data[sl], self.labels, dom)
            lv = sa_geometry_levels(stack, pos)
            if levels is None:
                levels = [{'idx': [], 'nbr': [], 'newpos': [], 'n': v['n']} for v in lv]
            b = pos.shape[0]
            for acc, v in zip(levels, lv):
                m = v['idx'].shape[1]
                base = (torch.arange(b, device=pos.device) * v['n']).view(b, 1)
                acc['idx'].append(v['idx'] - base)
                nbr = v['nbr'].view(b, m, -1)
                acc['nbr'].append(torch.where(nbr >= 0, nbr - base.view(b, 1, 1).to(torch.int32), nbr))
                acc['newpos'].append(v['newpos'])
        self.geo_cache = [{'idx': torch.cat(a['idx']).contiguous(), 'nbr': torch.cat(a['nbr']).contiguous(),
                           'newpos': torch.cat(a['newpos']).contiguous(), 'n': a['n']} for a in levels]
        self.geo_cache_key = tuple((l.ratio, l.radius, l.max_neighbors) for l in stack.levels)

    def drop_geometry_cache(self) -> None:
        self.geo_cache = None

    def _cached_geometry(self, ids: Tensor) -> list:
        flat = []
        for lv in self.geo_cache:
            flat += [lv['idx'], lv['nbr'], lv['newpos']]
        got = ops.gather_blocks_multi(flat, ids)                                              # one launch
        geo = []
        for i, lv in enumerate(self.geo_cache):
            idx_local, nbr_local, newpos = got[3 * i:3 * i + 3]
            idx, slots = ops.sa_cached_geometry(idx_local, nbr_local, lv['n'])               # one launch per level
            geo.append({'idx': idx, 'slots': slots, 'newpos': newpos})
        return geo

    # ---- batches (reference: collate_fn) -----------------------------------------------------------------------
    def _gather(self, ids: Tensor) -> FoamData:
        names = list(self.domain)
        out = ops.gather_blocks_multi([self.data] + [self.domain[k] for k in names], ids)     # one launch
        fd = FoamData(out[0], self.labels, dict(zip(names, out[1:])))
        if self.geo_cache is not None:
            fd.geometry = self._cached_geometry(ids)
        return fd

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

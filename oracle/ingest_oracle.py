"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64 like the reference) of the per-case features and the
collation the reference performs on the host while building batches:

  add_sdf          FoamDataset.add_sdf          dataset/foam_dataset.py:360-380
  boundary_one_hot FoamDataset.add_boundary_id  dataset/foam_dataset.py:382-395
  collate          collate_fn                   dataset/foam_dataset.py:83-90

Pinned on the unmodified reference: tests/golden/make_ingest_golden.py runs FoamDataset.add_sdf / add_boundary_id on
seeded frames and stores inputs and outputs in tests/golden/ingest.npz; tests/test_oracle_golden.py re-checks this file
against those vectors.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.
"""
from __future__ import annotations

import numpy as np


def add_sdf(internal_pos: np.ndarray, boundary_pos: np.ndarray, cell_to_region: np.ndarray, coord_scale=None):
    """-> (sdf of the internal points, sdf of the boundary points).

    foam_dataset.py:366-373: distances are measured on de-normalised coordinates (inverse_transform of the 'C' scaler;
    only its per-axis factor survives in a difference) from ALL boundary points; :375-376 min over the targets, divided
    by the largest value over internal + boundary points; :378-381 sign (0.5 - cellToRegion) * 2 inside, + on the boundary.
    """
    pts = np.concatenate([internal_pos, boundary_pos]).astype(np.float64)
    tgt = np.asarray(boundary_pos, dtype=np.float64)
    if coord_scale is not None:
        pts = pts * np.asarray(coord_scale, dtype=np.float64)
        tgt = tgt * np.asarray(coord_scale, dtype=np.float64)
    best = np.full(len(pts), np.inf)
    for lo in range(0, len(tgt), 512):          # tiles keep the (n x tile) distance block small
        d = pts[:, None, :] - tgt[None, lo:lo + 512, :]
        best = np.minimum(best, np.sqrt((d * d).sum(-1)).min(axis=1))
    sdf = best / best.max()
    ni = len(internal_pos)
    sign = (0.5 - np.asarray(cell_to_region, dtype=np.float64).reshape(-1)) * 2
    return sdf[:ni] * sign, sdf[ni:]


def boundary_classes(names) -> tuple[list, np.ndarray]:
    """foam_dataset.py:391-393: OneHotEncoder orders its categories lexicographically -> (sorted names, class per row)."""
    cats = sorted(set(names))
    return cats, np.array([cats.index(n) for n in names], dtype=np.int32)


def boundary_one_hot(names, n_internal: int) -> np.ndarray:
    """-> [n_internal + len(names), n_classes]: zeros for the internal rows (:388-389), one-hot for the boundary rows."""
    cats, cls = boundary_classes(names)
    out = np.zeros((n_internal + len(names), len(cats)))
    out[n_internal + np.arange(len(names)), cls] = 1.0
    return out


def collate(datas, domains):
    """foam_dataset.py:87-90: stack the data tensors and every sub-domain's ids."""
    return np.stack(datas), {k: np.stack([d[k] for d in domains]) for k in domains[0]}

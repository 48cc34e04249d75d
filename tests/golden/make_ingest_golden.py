"""Golden vectors of the reference's host-side feature code (dataset/foam_dataset.py:360-395), produced by running the
UNMODIFIED FoamDataset.add_sdf / add_boundary_id (imported through oracle/ref_shim.py) in the build container:

    python tests/golden/make_ingest_golden.py      -> tests/golden/ingest.npz

Each case builds seeded internal / boundary frames with the column layout load_case hands to add_features (MultiIndex
columns ('C', axis), 'cellToRegion'; boundary rows indexed by patch name), calls the two methods on a stand-in `self`
that only carries `normalizers`, and stores inputs and the columns they added.  Coordinates are float32-representable
so that the float32 device path sees the same inputs.
"""
import os
import sys
import types

import numpy as np
import pandas

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ingest_oracle, ref_shim  # noqa: E402

CASES = {  # name: (dims, n_internal, patches {name: rows}, coordinate scaler, seed)
    'abc3d': (3, 700, {'inlet': 60, 'outlet': 50, 'walls': 160, 'interface': 130}, None, 1),
    'duct2d_minmax': (2, 400, {'walls': 90, 'inlet': 30, 'interface': 80, 'outlet': 33}, 'minmax', 2),
    'std3d_ragged': (3, 1300, {'b': 700, 'a': 411}, 'standard', 3),
}


def frames(dims, ni, patches, rng):
    axes = ['x', 'y', 'z'][:dims]
    cols = pandas.MultiIndex.from_tuples([('C', a) for a in axes] + [('cellToRegion', '')])
    ci = rng.random((ni, dims)).astype(np.float32).astype(np.float64)
    region = (rng.random(ni) < 0.3).astype(np.float64)
    internal = pandas.DataFrame(np.concatenate([ci, region[:, None]], 1), columns=cols,
                                index=pandas.Index(['internal'] * ni))
    names = sum([[k] * v for k, v in patches.items()], [])
    nb = len(names)
    cb = rng.random((nb, dims)).astype(np.float32).astype(np.float64)
    face = rng.integers(0, dims, nb)
    cb[np.arange(nb), face] = np.round(cb[np.arange(nb), face])          # boundary points sit on the box faces
    boundary = pandas.DataFrame(np.concatenate([cb, np.zeros((nb, 1))], 1), columns=cols, index=pandas.Index(names))
    boundary = boundary.sort_index(axis=0)                              # load_case: sample_boundary(...).sort_index(axis=0)
    return internal, boundary


def main():
    ref_shim.install()
    from dataset.foam_dataset import FoamDataset, Normalizer, StandardScaler
    out = {}
    for name, (dims, ni, patches, scaler, seed) in CASES.items():
        rng = np.random.default_rng(seed)
        internal, boundary = frames(dims, ni, patches, rng)
        normalizers, scale = {}, None
        if scaler == 'minmax':
            lo, hi = np.array([-1.0, 0.5])[:dims], np.array([3.0, 2.0])[:dims]
            normalizers['C'] = Normalizer(lo, hi)
            scale = hi - lo
        elif scaler == 'standard':
            std, mean = np.array([0.5, 2.0, 1.25])[:dims], np.array([0.1, -0.2, 0.3])[:dims]
            normalizers['C'] = StandardScaler(std, mean)
            scale = std
        this = types.SimpleNamespace(normalizers=normalizers)
        pos_i, pos_b = internal['C'].values.copy(), boundary['C'].values.copy()
        region = internal['cellToRegion'].values.flatten().copy()
        patch_names = list(boundary.index.values)
        FoamDataset.add_sdf(this, internal, boundary)
        FoamDataset.add_boundary_id(this, internal, boundary)
        sdf_i, sdf_b = internal['sdf'].values.flatten(), boundary['sdf'].values.flatten()
        ohe = np.concatenate([internal['boundaryId'].values, boundary['boundaryId'].values])
        # the restatement must agree with the reference before anything is written
        mine_i, mine_b = ingest_oracle.add_sdf(pos_i, pos_b, region, scale)
        assert np.allclose(mine_i, sdf_i, rtol=1e-12, atol=1e-15) and np.allclose(mine_b, sdf_b, rtol=1e-12, atol=1e-15), name
        assert np.array_equal(ingest_oracle.boundary_one_hot(patch_names, ni), ohe), name
        cats, cls = ingest_oracle.boundary_classes(patch_names)
        out[f'{name}/pos_internal'], out[f'{name}/pos_boundary'], out[f'{name}/region'] = pos_i, pos_b, region
        out[f'{name}/coord_scale'] = np.zeros(0) if scale is None else scale
        out[f'{name}/boundary_class'] = cls
        out[f'{name}/n_classes'] = np.array([len(cats)])
        out[f'{name}/sdf_internal'], out[f'{name}/sdf_boundary'], out[f'{name}/one_hot'] = sdf_i, sdf_b, ohe
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'ingest.npz'), **out)
    print('wrote ingest.npz:', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()

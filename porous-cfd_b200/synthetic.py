"""Synthetic FoamData batches and model specs for the BASELINE.json configs.

There is no OpenFOAM data (and no network) where this runs, so the bench, the smoke test and the
parity tests use synthetic point clouds with the exact tensor layout the reference's dataset
produces (reference dataset/foam_dataset.py:296-313, 335-349, 406-437):

  * data   (B, N, F) fp32, internal rows first, then boundary rows grouped by boundary name;
  * labels ordered dict, single-column labels first (dict position == column), then multi labels;
  * domain {name: (B, n) int64 row ids}: internal, boundary, one range per named boundary, obs.

The layer shapes are the ones the reference's example scripts pass (examples/abc/train.py:26-49,
examples/duct_variable_boundary/train.py:28-37, examples/windbreaks/train.py:39-52,
examples/manufactured_solutions/train.py:13-27).  This module has no CUDA dependency.
"""
from __future__ import annotations

import copy

import torch

DIM_NAMES = ['x', 'y', 'z']

# ---- dataset layouts -------------------------------------------------------------------------

LAYOUTS = {
    # name: (dims, boundary names (sorted, as pandas sort_index leaves them), column groups)
    'abc': dict(dims=3, boundaries=['inlet', 'interface', 'outlet', 'walls'],
                columns=['C', 'U', 'p', 'cellToRegion', 'd', 'f', 'Ux-inlet', 'sdf', 'boundaryId']),
    'duct_variable': dict(dims=2, boundaries=['inlet', 'interface', 'outlet', 'walls'],
                          columns=['C', 'U', 'p', 'cellToRegion', 'd', 'f', 'U-inlet', 'sdf', 'boundaryId']),
    'windbreaks': dict(dims=3, boundaries=['ground', 'inlet', 'interface', 'outlet', 'top'],
                       columns=['C', 'U', 'p', 'cellToRegion', 'd', 'f', 'Ux-inlet', 'sdf', 'boundaryId']),
    'manufactured': dict(dims=2, boundaries=['interface', 'walls'],
                         columns=['C', 'cellToRegion', 'sdf', 'boundaryId', 'f', 'U', 'p']),
}
VECTOR_FIELDS = {'C', 'U', 'd', 'f', 'U-inlet'}


def build_labels(layout: str) -> dict:
    """Label dict in the order reference get_labels() emits it (dataset/foam_dataset.py:296-313)."""
    lay = LAYOUTS[layout]
    dims = lay['dims']
    singles, multis = {}, {}
    for col in lay['columns']:
        if col in VECTOR_FIELDS:
            subs = [col + DIM_NAMES[i] for i in range(dims)]
        elif col == 'boundaryId':
            subs = ['boundaryId' + b for b in lay['boundaries']]
        else:
            singles[col] = None
            continue
        for s in subs:
            singles[s] = None
        multis[col] = subs
    return {**singles, **multis}


def n_features(layout: str) -> int:
    return sum(1 for v in build_labels(layout).values() if v is None)


def make_batch(layout: str, n_geometries: int, n_internal: int, n_boundary: int, n_obs: int,
               seed: int = 8421, device='cpu'):
    """(data, labels, domain) of one collated batch, deterministic in `seed`."""
    lay = LAYOUTS[layout]
    dims, bnames = lay['dims'], lay['boundaries']
    labels = build_labels(layout)
    keys = list(labels.keys())
    col = lambda name: keys.index(name)
    g = torch.Generator().manual_seed(seed)
    B, NI, NB = n_geometries, n_internal, n_boundary
    N = NI + NB
    F = n_features(layout)
    data = torch.zeros(B, N, F)

    def put(name, values):  # values (B, N, k)
        subs = labels[name] if labels[name] else [name]
        for i, s in enumerate(subs):
            data[..., col(s)] = values[..., i]

    coords = torch.rand(B, N, dims, generator=g) * 2 - 1
    put('C', coords)
    zone = torch.zeros(B, N, 1)
    zone[:, :NI, 0] = (torch.rand(B, NI, generator=g) < 0.3).float()
    put('cellToRegion', zone)
    sdf = torch.rand(B, N, 1, generator=g) * 2 - 1
    sdf[:, NI:] = sdf[:, NI:].abs()          # boundary points are positive (foam_dataset.py:378-381)
    put('sdf', sdf)
    # one-hot boundary id on boundary rows, equal contiguous ranges per boundary name
    nbn = len(bnames)
    edges = [NI + (NB * i) // nbn for i in range(nbn + 1)]
    bid = torch.zeros(B, N, nbn)
    for i in range(nbn):
        bid[:, edges[i]:edges[i + 1], i] = 1.0
    put('boundaryId', bid)
    if layout == 'manufactured':
        # analytic fields of examples/manufactured_solutions/manufactured_dataset.py:46-67 (d=50, f=1, nu=0.01)
        x, y = coords[..., 0], coords[..., 1]
        ux, uy = torch.sin(y) * torch.cos(x), -torch.sin(x) * torch.cos(y)
        pr = -0.25 * (torch.cos(2 * x) + torch.cos(2 * y))
        mag = torch.sqrt(ux ** 2 + uy ** 2)
        z = zone[..., 0]
        fx = 2 * 0.01 * torch.cos(x) * torch.sin(y) + (0.01 * 50 + 0.5 * 1 * mag) * ux * z
        fy = -2 * 0.01 * torch.sin(x) * torch.cos(y) + (0.01 * 50 + 0.5 * 1 * mag) * uy * z
        put('f', torch.stack([fx, fy], -1))
        put('U', torch.stack([ux, uy], -1))
        put('p', pr[..., None])
    else:
        put('U', torch.randn(B, N, dims, generator=g))
        put('p', torch.randn(B, N, 1, generator=g))
        put('d', torch.rand(B, N, dims, generator=g) * zone)
        put('f', torch.rand(B, N, dims, generator=g) * zone)
        inlet = bnames.index('inlet')
        if 'U-inlet' in lay['columns']:
            v = torch.zeros(B, N, dims)
            v[:, edges[inlet]:edges[inlet + 1]] = torch.randn(B, edges[inlet + 1] - edges[inlet], dims, generator=g)
            put('U-inlet', v)
        else:
            v = torch.zeros(B, N, 1)
            v[:, edges[inlet]:edges[inlet + 1]] = torch.randn(B, edges[inlet + 1] - edges[inlet], 1, generator=g)
            put('Ux-inlet', v)

    domain = {'internal': torch.arange(NI).repeat(B, 1), 'boundary': (NI + torch.arange(NB)).repeat(B, 1)}
    for i, b in enumerate(bnames):
        domain[b] = torch.arange(edges[i], edges[i + 1]).repeat(B, 1)
    if n_obs > 0:
        domain['obs'] = torch.stack([torch.randperm(NI, generator=g)[:n_obs] for _ in range(B)])
    else:
        domain['obs'] = torch.zeros(B, 0, dtype=torch.int64)
    domain = {k: v.to(torch.int64).contiguous().to(device) for k, v in domain.items()}
    return data.to(device), labels, domain


def make_scalers(dims: int) -> dict:
    """Fixed literal scaler statistics."""
    f = lambda *v: torch.tensor(v, dtype=torch.float32)
    return {
        'C_std': f(0.5, 0.2, 0.2)[:dims], 'C_mean': f(0.1, -0.05, 0.02)[:dims],
        'U_std': f(0.1, 0.05, 0.05)[:dims], 'U_mean': f(0.3, 0.01, -0.02)[:dims],
        'p_std': f(0.02), 'p_mean': f(0.005),
        'd_min': f(0.0, 0.0, 0.0)[:dims], 'd_max': f(30000.0, 20000.0, 10000.0)[:dims],
        'f_min': f(0.0, 0.0, 0.0)[:dims], 'f_max': f(80.0, 60.0, 40.0)[:dims],
    }


# ---- model specs -----------------------------------------------------------------------------

def _weights(dims, continuity, momentum, boundary, observations):
    return [continuity] + [momentum] * dims + [boundary] * (dims + 1) + [observations] * (dims + 1)


def model_spec(name: str) -> dict:
    """Spec dict of a named configuration (consumed by the oracle, the golden generator and the
    host-side model factory alike)."""
    s = copy.deepcopy(_SPECS[name])
    s['scalers'] = make_scalers(s['dims']) if s['loss'] != 'manufactured' else None
    return s


_SPECS = {
    # BASELINE config 1: examples/abc/train.py:26-34
    'abc_pipn': dict(kind='PipnFoam', layout='abc', dims=3, activation='silu', loss='fixed',
                     nu=1489.4e-6, d=30000.0, f=79.731, enable_data_loss=True,
                     fe_local_layers=[3, 64, 64], fe_global_layers=[69, 96, 128, 1024],
                     seg_layers=[1088, 512, 256, 128, 4], seg_dropout=[0.03, 0.02, 0, 0],
                     loss_weights=_weights(3, 1, 1, 1, 100)),
    # BASELINE config 2: examples/abc/train.py:36-49
    'abc_pipn_pp': dict(kind='PipnFoamPp', layout='abc', dims=3, activation='silu', loss='fixed',
                        nu=1489.4e-6, d=30000.0, f=79.731, enable_data_loss=True,
                        fe_local_layers=[3, 64, 64], seg_layers=[1088, 384, 128, 4], seg_dropout=[0.03, 0, 0],
                        fe_radius=[0.5, 1], fe_fraction=[0.5, 0.25],
                        fe_global_layers=[[10, 64, 128], [131, 128, 256], [259, 256, 1024]],
                        max_neighbors=16, loss_weights=_weights(3, 1, 1, 1, 100)),
    # BASELINE config 3: examples/duct_variable_boundary/train.py:28-37
    'duct_pigano': dict(kind='PiGano', layout='duct_variable', dims=2, activation='silu', loss='variable',
                        nu=1489.4e-6, enable_data_loss=True, out_features=3,
                        branch_layers=[8, 128, 352, 352, 352], geometry_layers=[7, 64, 176, 176, 176],
                        local_layers=[2, 64, 176, 176, 176], n_operators=4, operator_dropout=[0, 0.1, 0.1, 0],
                        variable_boundaries={'Subdomains': ['inlet', 'internal'], 'Features': ['U-inlet', 'd', 'f']},
                        loss_weights=_weights(2, 1, 1, 1, 100)),
    # BASELINE config 4: examples/windbreaks/train.py:39-52
    'windbreaks_pigano_pp': dict(kind='PiGanoPp', layout='windbreaks', dims=3, activation='silu', loss='variable',
                                 nu=14.61e-6, enable_data_loss=True, out_features=4,
                                 branch_layers=[10, 256, 256, 512],
                                 geometry_layers=[[11, 64, 128], [131, 128], [131, 256, 256]],
                                 geometry_radius=[0.5, 1], geometry_fraction=[0.5, 0.25],
                                 local_layers=[3, 256, 256, 256], n_operators=4,
                                 operator_dropout=[0, 0.15, 0.15, 0], max_neighbors=64,
                                 variable_boundaries={'Subdomains': ['inlet', 'internal'],
                                                      'Features': ['Ux-inlet', 'd', 'f']},
                                 loss_weights=_weights(3, 10, 10, 1, 1)),
    # BASELINE config 5: examples/manufactured_solutions/train.py:19-27
    'manufactured_pipn_pp': dict(kind='PipnManufacturedPorousPp', layout='manufactured', dims=2, activation='tanh',
                                 loss='manufactured', nu=0.01, d=50.0, f=1.0, enable_data_loss=False,
                                 fe_local_layers=[2, 64, 64],
                                 fe_global_layers=[[6, 64], [66, 128], [130, 1024]],
                                 fe_radius=[0.6, 1.2], fe_fraction=[0.5, 0.25],
                                 seg_layers=[1088, 512, 256, 128, 3], seg_dropout=None, max_neighbors=64,
                                 loss_weights=None),
    # examples/manufactured_solutions/train.py:13-17 (vanilla PIPN; max-pool coupling)
    'manufactured_pipn': dict(kind='PipnManufactured', layout='manufactured', dims=2, activation='tanh',
                              fe_activation='tanh', loss='manufactured', nu=0.01, d=50.0, f=1.0,
                              enable_data_loss=False, fe_local_layers=[2, 64, 64],
                              fe_global_layers=[67, 64, 128, 1024], seg_layers=[1088, 512, 256, 128, 3],
                              seg_dropout=None, loss_weights=None),
    # ---- reduced-width variants of the same architectures for fast CPU parity fixtures ----
    'tiny_pipn': dict(kind='PipnFoam', layout='abc', dims=3, activation='silu', loss='fixed',
                      nu=1489.4e-6, d=30000.0, f=79.731, enable_data_loss=True,
                      fe_local_layers=[3, 16, 16], fe_global_layers=[21, 24, 32],
                      seg_layers=[48, 24, 16, 4], seg_dropout=[0.03, 0, 0],
                      loss_weights=_weights(3, 1, 1, 1, 100)),
    'tiny_pipn_pp': dict(kind='PipnFoamPp', layout='abc', dims=3, activation='silu', loss='fixed',
                         nu=1489.4e-6, d=30000.0, f=79.731, enable_data_loss=True,
                         fe_local_layers=[3, 16, 16], seg_layers=[48, 24, 16, 4], seg_dropout=[0.03, 0, 0],
                         fe_radius=[0.5, 1], fe_fraction=[0.5, 0.25],
                         fe_global_layers=[[10, 16, 24], [27, 24, 32], [35, 32, 32]],
                         max_neighbors=8, loss_weights=_weights(3, 1, 1, 1, 100)),
    'tiny_pigano': dict(kind='PiGano', layout='duct_variable', dims=2, activation='silu', loss='variable',
                        nu=1489.4e-6, enable_data_loss=True, out_features=3,
                        branch_layers=[8, 16, 40], geometry_layers=[7, 16, 20], local_layers=[2, 16, 20],
                        n_operators=3, operator_dropout=[0, 0.1, 0],
                        variable_boundaries={'Subdomains': ['inlet', 'internal'], 'Features': ['U-inlet', 'd', 'f']},
                        loss_weights=_weights(2, 1, 1, 1, 100)),
    'tiny_pigano_pp': dict(kind='PiGanoPp', layout='windbreaks', dims=3, activation='silu', loss='variable',
                           nu=14.61e-6, enable_data_loss=True, out_features=4, branch_layers=[10, 16, 40],
                           geometry_layers=[[11, 16, 24], [27, 24], [27, 24, 24]],
                           geometry_radius=[0.5, 1], geometry_fraction=[0.5, 0.25],
                           local_layers=[3, 16, 16], n_operators=2, operator_dropout=[0, 0.15], max_neighbors=8,
                           variable_boundaries={'Subdomains': ['inlet', 'internal'],
                                                'Features': ['Ux-inlet', 'd', 'f']},
                           loss_weights=_weights(3, 10, 10, 1, 1)),
    'tiny_manufactured_pp': dict(kind='PipnManufacturedPorousPp', layout='manufactured', dims=2, activation='tanh',
                                 loss='manufactured', nu=0.01, d=50.0, f=1.0, enable_data_loss=False,
                                 fe_local_layers=[2, 16, 16], fe_global_layers=[[6, 16], [18, 24], [26, 32]],
                                 fe_radius=[0.6, 1.2], fe_fraction=[0.5, 0.25],
                                 seg_layers=[48, 24, 16, 3], seg_dropout=None, max_neighbors=64,
                                 loss_weights=None),
    'tiny_manufactured': dict(kind='PipnManufactured', layout='manufactured', dims=2, activation='tanh',
                              fe_activation='tanh', loss='manufactured', nu=0.01, d=50.0, f=1.0,
                              enable_data_loss=False, fe_local_layers=[2, 16, 16],
                              fe_global_layers=[19, 24, 32], seg_layers=[48, 24, 16, 3],
                              seg_dropout=None, loss_weights=None),
}

SPEC_NAMES = tuple(_SPECS)


def rescale_weights(params: dict, gain: float = 3.0, seed: int = 3) -> dict:
    """Scale every weight MATRIX by `gain` (biases untouched) so that derivative-dependent loss
    terms are O(1e-3..1e2) instead of vanishing at default init."""
    out = {}
    for k, v in params.items():
        out[k] = v * gain if v.dim() == 2 else v.clone()
    return out

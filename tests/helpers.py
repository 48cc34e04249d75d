"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TINY = ['tiny_pipn', 'tiny_pipn_pp', 'tiny_pigano', 'tiny_pigano_pp', 'tiny_manufactured_pp', 'tiny_manufactured']
TINY_SHAPE = dict(n_geometries=2, n_internal=40, n_boundary=24, n_obs=10)


def load_fixture(name):
    """-> (data, domain, params, outputs{mode: {...}}) as torch tensors."""
    z = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    data = torch.from_numpy(z['data'])
    domain = {k[len('domain/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('domain/')}
    params = {k[len('param/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('param/')}
    out = {}
    for mode in ('reference', 'true'):
        out[mode] = {'loss': torch.from_numpy(z[f'{mode}/loss']), 'losses': torch.from_numpy(z[f'{mode}/losses']),
                     'u_error': torch.from_numpy(z[f'{mode}/u_error']), 'p_error': torch.from_numpy(z[f'{mode}/p_error']),
                     'grads': {k[len(f'{mode}/grad/'):]: torch.from_numpy(z[k]) for k in z.files
                               if k.startswith(f'{mode}/grad/')}}
    out['seed'] = int(z['seed'])
    return data, domain, params, out


def flat(grads: dict, keys) -> torch.Tensor:
    return torch.cat([grads[k].detach().flatten().double().cpu() for k in keys])


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(((a - b).abs() / (b.abs() + 1e-30)).max())

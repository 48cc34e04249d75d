"""Builds a model from a spec dict (porous_cfd_b200.synthetic.model_spec) by calling the model
constructors with the same keyword arguments the reference's example scripts use
(examples/abc/train.py:19-80, examples/duct_variable_boundary/train.py:21-83,
examples/windbreaks/train.py:21-71, examples/manufactured_solutions/train.py:9-29)."""
from __future__ import annotations

import numpy as np
from torch.nn import SiLU, Tanh

from .dataset.foam_dataset import Normalizer, StandardScaler
from .models.losses import FixedLossScaler
from .models.pi_gano.pi_gano import PiGano
from .models.pi_gano.pi_gano_pp import PiGanoPp
from .models.pipn.pipn_baseline import PipnManufactured, PipnManufacturedPorousPp
from .models.pipn.pipn_foam import PipnFoam, PipnFoamPp

ACTIVATIONS = {'silu': SiLU, 'tanh': Tanh}


def build_scalers(spec: dict):
    sc = spec['scalers']
    f64 = lambda t: t.numpy().astype(np.float64)
    return {'C': StandardScaler(f64(sc['C_std']), f64(sc['C_mean'])),
            'U': StandardScaler(f64(sc['U_std']), f64(sc['U_mean'])),
            'p': StandardScaler(f64(sc['p_std']), f64(sc['p_mean'])),
            'd': Normalizer(f64(sc['d_min']), f64(sc['d_max'])),
            'f': Normalizer(f64(sc['f_min']), f64(sc['f_max']))}


def build_loss_scaler(spec: dict):
    w, d = spec.get('loss_weights'), spec['dims']
    if w is None:
        return None
    return FixedLossScaler({'continuity': w[:1], 'momentum': w[1:1 + d], 'boundary': w[1 + d:2 + 2 * d],
                            'observations': w[2 + 2 * d:]})


def build_model(spec: dict):
    act = ACTIVATIONS[spec['activation']]
    kind = spec['kind']
    if spec['loss'] != 'manufactured':
        scalers, loss_scaler = build_scalers(spec), build_loss_scaler(spec)
    if kind == 'PipnFoam':
        return PipnFoam(nu=spec['nu'], d=spec['d'], f=spec['f'], fe_local_layers=spec['fe_local_layers'],
                        fe_global_layers=spec['fe_global_layers'], seg_layers=spec['seg_layers'],
                        seg_dropout=spec['seg_dropout'], scalers=scalers, loss_scaler=loss_scaler, activation=act)
    if kind == 'PipnFoamPp':
        return PipnFoamPp(nu=spec['nu'], d=spec['d'], f=spec['f'], fe_local_layers=spec['fe_local_layers'],
                          seg_layers=spec['seg_layers'], seg_dropout=spec['seg_dropout'], fe_radius=spec['fe_radius'],
                          fe_fraction=spec['fe_fraction'], fe_global_layers=spec['fe_global_layers'], scalers=scalers,
                          loss_scaler=loss_scaler, max_neighbors=spec['max_neighbors'], activation=act)
    if kind == 'PipnManufactured':
        return PipnManufactured(nu=spec['nu'], d=spec['d'], f=spec['f'], fe_local_layers=spec['fe_local_layers'],
                                fe_global_layers=spec['fe_global_layers'], seg_layers=spec['seg_layers'],
                                activation=act)
    if kind == 'PipnManufacturedPorousPp':
        return PipnManufacturedPorousPp(nu=spec['nu'], d=spec['d'], f=spec['f'],
                                        fe_local_layers=spec['fe_local_layers'],
                                        fe_global_layers=spec['fe_global_layers'],
                                        fe_global_radius=spec['fe_radius'], fe_global_fraction=spec['fe_fraction'],
                                        seg_layers=spec['seg_layers'], activation=act)
    if kind == 'PiGano':
        return PiGano(nu=spec['nu'], out_features=spec['out_features'], branch_layers=spec['branch_layers'],
                      geometry_layers=spec['geometry_layers'], local_layers=spec['local_layers'],
                      n_operators=spec['n_operators'], operator_dropout=spec['operator_dropout'], scalers=scalers,
                      variable_boundaries=spec['variable_boundaries'], loss_scaler=loss_scaler, activation=act)
    if kind == 'PiGanoPp':
        return PiGanoPp(nu=spec['nu'], out_features=spec['out_features'], branch_layers=spec['branch_layers'],
                        geometry_layers=spec['geometry_layers'], geometry_radius=spec['geometry_radius'],
                        geometry_fraction=spec['geometry_fraction'], local_layers=spec['local_layers'],
                        n_operators=spec['n_operators'], operator_dropout=spec['operator_dropout'], scalers=scalers,
                        variable_boundaries=spec['variable_boundaries'], loss_scaler=loss_scaler, activation=act,
                        max_neighbors=spec['max_neighbors'])
    raise KeyError(kind)

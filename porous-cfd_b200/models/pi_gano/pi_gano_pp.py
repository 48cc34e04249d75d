"""PI-GANO++: set-abstraction geometry encoder over the boundary points
(reference models/pi_gano/pi_gano_pp.py:13-87)."""
from __future__ import annotations

import torch
from torch.nn import Linear, Module, SiLU
from torch.optim.lr_scheduler import ExponentialLR

from ..losses import LossScaler
from ..modules import MLP, Branch, GeometryEncoderPp, NeuralOperatorSequential
from .base import PiGanoBase


class PiGanoPp(PiGanoBase):
    def __init__(self, nu: float, out_features: int, branch_layers: list[int], geometry_layers: list[list[int]],
                 geometry_radius: list[int], geometry_fraction: list[float], local_layers: list[int],
                 n_operators: int, operator_dropout: list[float], scalers: dict, variable_boundaries: dict[str, list],
                 loss_scaler: LossScaler = None, activation: type[Module] = SiLU, max_neighbors=64):
        super().__init__(nu, out_features, scalers, loss_scaler, variable_boundaries)
        self.branch = Branch(branch_layers, activation)
        self.geometry_encoder = GeometryEncoderPp(geometry_fraction, geometry_radius, geometry_layers, activation,
                                                  max_neighbors)
        self.points_encoder = MLP(local_layers, None, activation)
        width = geometry_layers[-1][-1] + local_layers[-1]
        self.neural_ops = NeuralOperatorSequential(n_operators, width, operator_dropout, activation)
        self.reduction = Linear(width, out_features)

    def build_plan(self) -> dict:
        plan = self.operator_plan('pigano_pp')
        plan['sa_stack'] = self.geometry_encoder.set_abstraction.module.stack()
        plan['geom_feature_order'] = ['C', 'boundaryId']      # reference models/pi_gano/pi_gano_pp.py:71
        return plan

    def configure_optimizers(self):
        optimizer = torch.optim.Adam(self.parameters(), lr=0.001)
        return [optimizer], [{'scheduler': ExponentialLR(optimizer, 0.999), 'interval': 'epoch'}]

"""PI-GANO (reference models/pi_gano/pi_gano.py:11-74)."""
from __future__ import annotations

import torch
from torch.nn import Linear, Module, SiLU
from torch.optim.lr_scheduler import ExponentialLR

from ..losses import LossScaler
from ..modules import MLP, Branch, GeometryEncoder, NeuralOperatorSequential
from .base import PiGanoBase


class PiGano(PiGanoBase):
    def __init__(self, nu: float, out_features: int, branch_layers: list[int], geometry_layers: list[int],
                 local_layers: list[int], n_operators: int, operator_dropout: list[float], scalers: dict,
                 variable_boundaries: dict[str, list], loss_scaler: LossScaler = None,
                 activation: type[Module] = SiLU):
        super().__init__(nu, out_features, scalers, loss_scaler, variable_boundaries)
        self.branch = Branch(branch_layers, activation)
        self.geometry_encoder = GeometryEncoder(geometry_layers, activation)
        self.points_encoder = MLP(local_layers, None, activation)
        width = geometry_layers[-1] + local_layers[-1]
        self.neural_ops = NeuralOperatorSequential(n_operators, width, operator_dropout, activation)
        self.reduction = Linear(width, out_features)

    def build_plan(self) -> dict:
        plan = self.operator_plan('pigano')
        geom, (gpend, _) = self.geometry_encoder.linear.chain()
        plan.update({'geom_layers': geom, 'geom_pending_act': gpend})
        return plan

    def configure_optimizers(self):
        optimizer = torch.optim.Adam(self.parameters(), lr=0.001)
        return [optimizer], [{'scheduler': ExponentialLR(optimizer, 0.999), 'interval': 'epoch'}]

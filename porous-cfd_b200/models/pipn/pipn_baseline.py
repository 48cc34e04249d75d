"""PIPN / PIPN++ for the manufactured-solution experiment: raw outputs, no data loss, Tanh
(reference models/pipn/pipn_baseline.py:12-124)."""
from __future__ import annotations

import torch
from torch.nn import Module, Tanh
from torch.optim.lr_scheduler import ExponentialLR

from ..losses import ContinuityLoss, MomentumLossManufactured
from ..model_base import PorousPinnBase
from ..modules import MLP, PointNetFeatureExtract, PointNetFeatureExtractPp
from .pipn_foam import pointnet_plan


class _ManufacturedBase(PorousPinnBase):
    def _init_losses(self, nu, d, f):
        self.momentum_loss = MomentumLossManufactured(nu, d, f)
        self.continuity_loss = ContinuityLoss()

    def loss_spec(self) -> dict:
        m = self.momentum_loss
        return {'kind': 'manufactured', 'nu': m.nu, 'd': m.d, 'f': m.f}

    def configure_optimizers(self):
        optimizer = torch.optim.Adam(self.parameters(), lr=0.001, eps=1e-6)
        return [optimizer], [{'scheduler': ExponentialLR(optimizer, 0.9995), 'interval': 'epoch'}]


class PipnManufactured(_ManufacturedBase):
    """Vanilla PIPN; like the reference, the feature extractor keeps its default Tanh whatever
    `activation` is (pipn_baseline.py:39).  Max-pool coupling in the Jacobian: see PipnFoam."""

    def __init__(self, nu: float, d: float, f: float, fe_local_layers: list[int], fe_global_layers: list[int],
                 seg_layers: list[int], activation: type[Module] = Tanh):
        super().__init__(seg_layers[-1], False, None)
        self.save_hyperparameters()
        self.feature_extract = PointNetFeatureExtract(fe_local_layers, fe_global_layers)
        self.decoder = MLP(seg_layers, None, activation, False)
        self._init_losses(nu, d, f)

    def build_plan(self) -> dict:
        return pointnet_plan(self, 'pipn')


class PipnManufacturedPorousPp(_ManufacturedBase):
    def __init__(self, nu: float, d: float, f: float, fe_local_layers: list[int], fe_global_layers: list[list[int]],
                 fe_global_radius: list[float], fe_global_fraction: list[float], seg_layers: list[int],
                 activation=Tanh):
        super().__init__(seg_layers[-1], False, None)
        self.save_hyperparameters()
        self.feature_extract = PointNetFeatureExtractPp(fe_local_layers, fe_global_layers, fe_global_fraction,
                                                        fe_global_radius, activation)
        self.decoder = MLP(seg_layers, None, activation, False)
        self._init_losses(nu, d, f)

    def build_plan(self) -> dict:
        # geometry features = [boundaryId, C]: the order differs from PipnFoamPp (pipn_baseline.py:110)
        return pointnet_plan(self, 'pipn_pp', ['boundaryId', 'C'])

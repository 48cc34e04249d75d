"""PIPN and PIPN++ with feature scaling, dropout and data loss (reference models/pipn/pipn_foam.py:15-166)."""
from __future__ import annotations

import torch
from torch.nn import Module, SiLU
from torch.optim.lr_scheduler import ExponentialLR

from ...dataset.foam_dataset import StandardScaler
from ...engine import ChainLayer
from ..losses import ContinuityLossStandardized, LossScaler, MomentumLossFixed
from ..model_base import PorousPinnBase
from ..modules import MLP, PointNetFeatureExtract, PointNetFeatureExtractPp, activation_name


def pointnet_plan(model, family: str, geom_feature_order=None) -> dict:
    """Executor plan shared by the PointNet-style models: local MLP -> [local | pooled global] -> decoder.
    The decoder's first Linear is split into its per-point column block (chain layer with a
    per-geometry constant) and its global column block (`concat_layer`), so the (B, N, 1024)
    broadcast of the reference (models/pipn/pipn_foam.py:96-97) is never materialised."""
    fe = model.feature_extract
    local, (pend_act, _) = fe.local_feature.chain()
    lw = local[-1].n
    dec = model.decoder.linears()
    first = dec[0]
    g_width = first.in_features - lw
    drop = model.decoder.dropout_p
    rest, _ = MLP.chain(model.decoder)
    point_layers = list(local)
    point_layers.append(ChainLayer(first.weight, None, 0, lw, first.out_features, act=pend_act, cvec_key='concat'))
    point_layers += rest[1:]
    plan = {'family': family, 'dims': model.dims, 'point_layers': point_layers,
            'concat_layer': ChainLayer(first.weight, first.bias, lw, g_width, first.out_features)}
    if family == 'pipn_pp':
        plan['sa_stack'] = fe.global_feature.module.stack()
        plan['geom_feature_order'] = geom_feature_order
    else:
        glayers, (gpend, _) = fe.global_feature.chain(first_act=pend_act)
        glayers[0].act_cols = lw          # [act(local) | boundaryId, sdf]: only the local block is activated
        plan.update({'local_layers': local, 'global_layers': glayers, 'global_pending_act': gpend})
    return plan


class PipnFoamBase(PorousPinnBase):
    def __init__(self, nu: float, d: float, f: float, out_features: int, scalers: dict[str, StandardScaler],
                 loss_scaler: LossScaler = None):
        super().__init__(out_features, True, loss_scaler)
        self.save_hyperparameters()
        self.u_scaler, self.p_scaler, self.points_scaler = scalers['U'], scalers['p'], scalers['C']
        self.momentum_loss = MomentumLossFixed(nu, d, f, self.u_scaler, self.points_scaler, self.p_scaler)
        self.continuity_loss = ContinuityLossStandardized(self.u_scaler, self.points_scaler)

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        for sc in (self.u_scaler, self.p_scaler, self.points_scaler):
            sc.to(*args, **kwargs).to(torch.float)
        return self

    def postprocess_out(self, u, p):
        return self.u_scaler.inverse_transform(u), self.p_scaler.inverse_transform(p)

    def loss_spec(self) -> dict:
        m = self.momentum_loss
        return {'kind': 'fixed', 'nu': m.nu, 'd': m.d, 'f': m.f, 'C': self.points_scaler, 'U': self.u_scaler,
                'p': self.p_scaler}

    def configure_optimizers(self):
        optimizer = torch.optim.Adam(self.parameters(), lr=0.001)
        return [optimizer], [{'scheduler': ExponentialLR(optimizer, 0.999), 'interval': 'epoch'}]


class PipnFoam(PipnFoamBase):
    """Vanilla PIPN.  NOTE: the reference's Jacobian for this model contains max-pool cross-point
    terms that the forward-mode jet does not carry yet ('next' row 1)."""

    def __init__(self, nu: float, d: float, f: float, fe_local_layers: list[int], fe_global_layers: list[int],
                 seg_layers: list[int], scalers: dict[str, StandardScaler], loss_scaler: LossScaler = None,
                 seg_dropout: list[float] = None, activation: type[Module] = SiLU):
        super().__init__(nu, d, f, seg_layers[-1], scalers, loss_scaler)
        self.feature_extract = PointNetFeatureExtract(fe_local_layers, fe_global_layers, activation)
        self.decoder = MLP(seg_layers, seg_dropout, activation, False)

    def build_plan(self) -> dict:
        return pointnet_plan(self, 'pipn')


class PipnFoamPp(PipnFoamBase):
    """PIPN++: set-abstraction geometry encoder over the boundary points."""

    def __init__(self, nu: float, d: float, f: float, fe_local_layers: list[int], fe_global_layers: list[list[int]],
                 fe_radius, fe_fraction, seg_layers: list[int], scalers: dict[str, StandardScaler],
                 loss_scaler: LossScaler = None, seg_dropout: list[float] = None, activation: type[Module] = SiLU,
                 max_neighbors=64):
        super().__init__(nu, d, f, seg_layers[-1], scalers, loss_scaler)
        self.feature_extract = PointNetFeatureExtractPp(fe_local_layers, fe_global_layers, fe_fraction, fe_radius,
                                                        activation, max_neighbors)
        self.decoder = MLP(seg_layers, seg_dropout, activation, False)

    def build_plan(self) -> dict:
        # geometry features = [C, boundaryId] (reference models/pipn/pipn_foam.py:154)
        return pointnet_plan(self, 'pipn_pp', ['C', 'boundaryId'])

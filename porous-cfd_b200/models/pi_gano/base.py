"""Common part of the PI-GANO models (reference models/pi_gano/base.py:9-76)."""
from __future__ import annotations

import torch

from ...engine import ChainLayer
from ..losses import ContinuityLossStandardized, LossScaler, MomentumLossVariable
from ..model_base import PorousPinnBase


class PiGanoBase(PorousPinnBase):
    def __init__(self, nu: float, out_features: int, scalers: dict, loss_scaler: LossScaler,
                 variable_boundaries: dict):
        super().__init__(out_features, True, loss_scaler)
        self.save_hyperparameters()
        self.u_scaler, self.p_scaler, self.points_scaler = scalers['U'], scalers['p'], scalers['C']
        self.d_scaler, self.f_scaler = scalers['d'], scalers['f']
        self.continuity_loss = ContinuityLossStandardized(self.u_scaler, self.points_scaler)
        self.momentum_loss = MomentumLossVariable(nu, self.u_scaler, self.points_scaler, self.p_scaler,
                                                  self.d_scaler, self.f_scaler)
        self.variable_boundaries = variable_boundaries

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        for sc in (self.u_scaler, self.p_scaler, self.points_scaler, self.d_scaler, self.f_scaler):
            sc.to(*args, **kwargs).to(torch.float)
        return self

    def postprocess_out(self, u, p):
        return self.u_scaler.inverse_transform(u), self.p_scaler.inverse_transform(p)

    def loss_spec(self) -> dict:
        return {'kind': 'variable', 'nu': self.momentum_loss.nu, 'C': self.points_scaler, 'U': self.u_scaler,
                'p': self.p_scaler, 'd_scaler': self.d_scaler, 'f_scaler': self.f_scaler}

    def operator_plan(self, family: str) -> dict:
        """points encoder -> neural operators (each multiplied by the branch embedding) -> reduction.
        Operator 0 consumes [local embedding | geometry embedding]: its geometry column block becomes
        the per-geometry constant (`concat_layer`), the (B, N, G) repeat of the reference
        (models/pi_gano/pi_gano.py:60-64) is never materialised."""
        enc, (pend_act, _) = self.points_encoder.chain()
        lw = enc[-1].n
        ops_ = self.neural_ops.operators()
        first = ops_[0].linear[0]
        g_width = first.in_features - lw
        layers = list(enc)
        layers.append(ChainLayer(first.weight, None, 0, lw, first.out_features, act=pend_act, cvec_key='concat'))
        prev = ops_[0]
        for op in ops_[1:]:
            lin = op.linear[0]
            layers.append(ChainLayer(lin.weight, lin.bias, 0, lin.in_features, lin.out_features, act=prev.act_name,
                                     drop_p=prev.drop_p, escale=True))
            prev = op
        red = self.reduction
        layers.append(ChainLayer(red.weight, red.bias, 0, red.in_features, red.out_features, act=prev.act_name,
                                 drop_p=prev.drop_p, escale=True))
        branch, (bpend, _) = self.branch.linear.chain()
        return {'family': family, 'dims': self.dims, 'point_layers': layers,
                'concat_layer': ChainLayer(first.weight, first.bias, lw, g_width, first.out_features),
                'branch_layers': branch, 'branch_pending_act': bpend,
                'variable_boundaries': self.variable_boundaries}

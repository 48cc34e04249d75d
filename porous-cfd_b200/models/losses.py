"""Residual-loss and loss-weighting modules with the reference's names and constructor arguments
(reference models/losses.py).  They hold the physics constants and scalers; the arithmetic
(continuity, momentum with Darcy-Forchheimer penalisation, MSE terms, weighting, reduction and the
gradient wrt the model output) is ONE fused CUDA pass, `pcfd_residual_loss` (csrc/residual.cu),
configured from these objects by `PorousPinnBase.residual_params`.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from ..dataset.foam_dataset import Normalizer, StandardScaler


class LossScaler(nn.Module):
    """Identity weighting (reference models/losses.py:23-36)."""

    def weights(self, n_terms: int) -> list[float]:
        return [1.0] * n_terms

    def forward(self, model, losses: Tensor) -> Tensor:
        return losses


class FixedLossScaler(LossScaler):
    """Fixed per-term weights in the order continuity, momentum, boundary[, observations]
    (reference models/losses.py:39-61).  The weights are applied inside the residual kernel."""

    def __init__(self, loss_weights: dict[str, list]):
        super().__init__()
        w = list(loss_weights['continuity']) + list(loss_weights['momentum']) + list(loss_weights['boundary'])
        if 'observations' in loss_weights:
            w += list(loss_weights['observations'])
        self.weights_t = torch.tensor(w, dtype=torch.float)

    def weights(self, n_terms: int) -> list[float]:
        w = [float(v) for v in self.weights_t.tolist()]
        if len(w) < n_terms:
            raise ValueError(f'FixedLossScaler holds {len(w)} weights, the step produces {n_terms} loss terms')
        return w[:n_terms]

    def forward(self, model, losses: Tensor) -> Tensor:
        return losses * self.weights_t.to(losses.device)

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        self.weights_t = self.weights_t.to(*args, **kwargs)
        return self


class RelobraloScaler(LossScaler):
    """ReLoBRaLo adaptive weighting (reference models/losses.py:64-124), same constructor, same registered
    buffers (`init_losses`, `prev_losses`, `lambda_ema`: reference checkpoints load with strict=True).

    The update runs on the device inside the fused step (`pcfd_relobralo_update`, csrc/residual.cu): the step
    evaluates the unscaled loss terms, updates the buffers and the weight vector in one single-thread kernel,
    then evaluates the weighted loss and its gradient with the device-resident weights -- no host read, so the
    whole step stays graph-capturable.  `global_step` is a device counter advanced by the kernel;
    `batch_size` is what the reference reads from `model.trainer.train_dataloader.batch_size` (set it with
    `set_batch_size`, default 1).  rho ~ Bernoulli(beta) comes from a counter-based hash of (seed, step), not from
    torch's global generator: identical in distribution, bit-identical to the reference only for beta in {0, 1}.
    """
    dynamic = True

    def __init__(self, num_losses: int, alpha=0.95, beta=0.99, tau=1.0, eps=1e-8):
        super().__init__()
        self.num_losses, self.alpha, self.beta, self.tau, self.eps = num_losses, alpha, beta, tau, eps
        self.register_buffer('init_losses', torch.zeros(num_losses))
        self.register_buffer('prev_losses', torch.zeros(num_losses))
        self.register_buffer('lambda_ema', torch.ones(num_losses))
        self.batch_size = 1
        self.seed = 8421
        self._step = None       # device int64 counter
        self._weights = None    # device float32[16]

    def set_batch_size(self, batch_size: int) -> None:
        self.batch_size = int(batch_size)

    def weights(self, n_terms: int) -> list[float]:
        return [1.0] * n_terms     # the static slot of the residual parameters; the live weights are on the device

    def device_state(self, device):
        if self._step is None or self._step.device != device:
            self._step = torch.zeros(1, dtype=torch.int64, device=device)
            self._weights = torch.ones(16, dtype=torch.float32, device=device)
        return self._step, self._weights

    def forward(self, model, losses: Tensor) -> Tensor:
        """Host-callable form (the fused step does not go through here): weights `losses` with the current state."""
        from .. import ops
        step, w = self.device_state(losses.device)
        ops.relobralo_update(losses.detach().contiguous(), self.num_losses, self.init_losses, self.prev_losses,
                             self.lambda_ema, step, self.batch_size, self.alpha, self.beta, self.tau, self.eps, self.seed, w)
        return w[:self.num_losses] * losses


class LossLogger:
    """Pairs loss labels with values and hands them to `module.log` (reference models/losses.py:127-146)."""

    def __init__(self, module, *loss_labels: str):
        self.loss_labels = loss_labels
        self.module = module

    def log(self, batch_size: int, *losses: Tensor):
        if len(losses) != len(self.loss_labels):
            print('Mismatching losses!')
        for label, value in zip(self.loss_labels, losses):
            self.module.log(label, value, on_step=False, on_epoch=True, batch_size=batch_size)


class _ResidualSpec(nn.Module):
    """Common holder: which residual variant the kernel evaluates and with which constants."""
    kind = 'fixed'

    def __init__(self):
        super().__init__()
        self.nu, self.d, self.f = 0.0, 0.0, 0.0
        self.u_scaler = self.points_scaler = self.p_scaler = self.d_scaler = self.f_scaler = None

    def func(self, *args):
        raise NotImplementedError('per-point residual fields come from PorousPinnBase.predict_step with '
                                  'verbose_predict=True (pcfd_residual_fields); the loss modules hold constants only')

    def forward(self, *args):
        raise NotImplementedError('losses are evaluated inside PorousPinnBase.training_step by pcfd_residual_loss')


class ContinuityLoss(_ResidualSpec):
    """div U = 0 on raw outputs (reference models/losses.py:149-164)."""
    kind = 'manufactured'


class ContinuityLossStandardized(_ResidualSpec):
    """div U = 0 on standardised outputs (reference models/losses.py:167-190)."""

    def __init__(self, u_scaler: StandardScaler, points_scaler: StandardScaler):
        super().__init__()
        self.u_scaler, self.points_scaler = u_scaler, points_scaler


class MomentumLossManufactured(_ResidualSpec):
    """Raw-output momentum residual with forcing term (reference models/losses.py:193-225)."""
    kind = 'manufactured'

    def __init__(self, nu: float, d: float, f: float):
        super().__init__()
        self.nu, self.d, self.f = nu, d, f


class MomentumLossFixed(_ResidualSpec):
    """Standardised outputs, scalar Darcy / Forchheimer coefficients (reference models/losses.py:228-270)."""
    kind = 'fixed'

    def __init__(self, nu: float, d: float, f: float, u_scaler: StandardScaler, points_scaler: StandardScaler,
                 p_scaler: StandardScaler):
        super().__init__()
        self.nu, self.d, self.f = nu, d, f
        self.u_scaler, self.points_scaler, self.p_scaler = u_scaler, points_scaler, p_scaler


class MomentumLossVariable(_ResidualSpec):
    """Standardised outputs, per-point per-component coefficients (reference models/losses.py:273-319)."""
    kind = 'variable'

    def __init__(self, nu: float, u_scaler: StandardScaler, points_scaler: StandardScaler, p_scaler: StandardScaler,
                 d_scaler: Normalizer, f_scaler: Normalizer):
        super().__init__()
        self.nu = nu
        self.u_scaler, self.points_scaler, self.p_scaler = u_scaler, points_scaler, p_scaler
        self.d_scaler, self.f_scaler = d_scaler, f_scaler

"""Residual-loss and loss-weighting modules with the reference's names and constructor arguments
(reference models/losses.py).  They hold the physics constants and scalers; the arithmetic
(continuity, momentum with Darcy-Forchheimer penalisation, MSE terms, weighting, reduction and the
gradient wrt the model output) is ONE fused CUDA pass, `pcfd_residual_loss` (csrc/residual.cu),
configured from these objects by `PorousPinnBase.residual_params`.
"""
from __future__ import annotations

import weakref

import torch
from torch import Tensor, nn

from ..dataset.foam_dataset import Normalizer, StandardScaler


class LossScaler(nn.Module):
    """Identity weighting (reference models/losses.py:23-36)."""

    def weight_list(self, n_terms: int) -> list[float]:
        """The static weights the residual kernel applies, one per loss term."""
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
        self.weights = torch.tensor(w, dtype=torch.float)      # tensor attribute, as in the reference (:53)

    def weight_list(self, n_terms: int) -> list[float]:
        w = [float(v) for v in self.weights.tolist()]
        if len(w) < n_terms:
            raise ValueError(f'FixedLossScaler holds {len(w)} weights, the step produces {n_terms} loss terms')
        return w[:n_terms]

    def forward(self, model, losses: Tensor) -> Tensor:
        return losses * self.weights.to(losses.device)

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        self.weights = self.weights.to(*args, **kwargs)
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
        self._resume_step = 0   # value the counter starts from when it is (re)created: > 0 after a state was loaded

    def set_batch_size(self, batch_size: int) -> None:
        self.batch_size = int(batch_size)

    def set_global_step(self, step: int) -> None:
        """The reference keys its update on `model.global_step`, which Lightning restores from the checkpoint; here the
        counter lives on the device, so a resumed run hands it over explicitly (common.training.train does)."""
        self._resume_step = int(step)
        if self._step is not None:
            self._step.fill_(int(step))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        # Buffers restored from a checkpoint belong to a run that is past its first step: the kernel's step-0 branch
        # would overwrite init_losses / prev_losses with the current losses.  Without an explicit global step the
        # counter restarts at 1 (any positive value keeps the restored statistics).
        key = prefix + 'init_losses'
        if key in state_dict and bool((state_dict[key] != 0).any()):
            self.set_global_step(max(1, self._resume_step))

    def weight_list(self, n_terms: int) -> list[float]:
        return [1.0] * n_terms     # the static slot of the residual parameters; the live weights are on the device

    def device_state(self, device):
        if self._step is None or self._step.device != device:
            self._step = torch.full((1,), self._resume_step, dtype=torch.int64, device=device)
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
        # weak: the model owns its loggers; a strong back-reference would make every model cyclic garbage that only
        # Python's cyclic collector can free (with its CUDA graphs and buffers, see engine.GraphedStep)
        self.module = weakref.proxy(module)

    def log(self, batch_size: int, *losses: Tensor):
        if len(losses) != len(self.loss_labels):
            print('Mismatching losses!')
        for label, value in zip(self.loss_labels, losses):
            self.module.log(label, value, on_step=False, on_epoch=True, batch_size=batch_size)


def _vec(x, n):
    t = torch.as_tensor(x, dtype=torch.float64).flatten().cpu()
    if t.numel() == 1:
        t = t.repeat(n)
    return [float(v) for v in t[:n]]


class _ResidualSpec(nn.Module):
    """Common part of the loss modules: the physics constants / scalers, and `func` / `forward` on explicit tensors
    through pcfd_residual_eval + pcfd_mean_squares (csrc/residual_eval.cu).  The training step does not call these: it
    evaluates residual, losses and gradient in the fused pcfd_residual_loss, configured from the same attributes by
    `PorousPinnBase.residual_params`."""
    kind = 'fixed'

    def __init__(self):
        super().__init__()
        self.nu, self.d, self.f = 0.0, 0.0, 0.0
        self.u_scaler = self.points_scaler = self.p_scaler = self.d_scaler = self.f_scaler = None

    def eval_params(self, dims: int):
        """pcfd_residual_params_t for the explicit-tensor kernel (column indices and weights unused)."""
        from .._lib import LOSS_KINDS, ResidualParams
        prm = ResidualParams()
        prm.dims, prm.loss_kind, prm.lap_mode, prm.enable_data_loss = dims, LOSS_KINDS[self.kind], 1, 0
        prm.nu, prm.d, prm.f = float(self.nu), float(self.d), float(self.f)
        ones, zeros = [1.0] * 3, [0.0] * 3
        prm.c_std[:], prm.u_std[:], prm.u_mean[:] = ones, ones, zeros
        prm.p_std, prm.p_mean = 1.0, 0.0
        if self.kind != 'manufactured':
            prm.c_std[:dims] = _vec(self.points_scaler.std, dims)
            prm.u_std[:dims], prm.u_mean[:dims] = _vec(self.u_scaler.std, dims), _vec(self.u_scaler.mean, dims)
            if self.p_scaler is not None:
                prm.p_std, prm.p_mean = _vec(self.p_scaler.std, 1)[0], _vec(self.p_scaler.mean, 1)[0]
        if self.kind == 'variable':
            prm.d_min[:dims], prm.d_range[:dims] = _vec(self.d_scaler.min, dims), _vec(self.d_scaler.range, dims)
            prm.f_min[:dims], prm.f_range[:dims] = _vec(self.f_scaler.min, dims), _vec(self.f_scaler.range, dims)
        return prm


class _Continuity(_ResidualSpec):
    def func(self, jacobian: Tensor) -> Tensor:
        """div = sum_i jac[..., i, i] * u_std_i / x_std_i  (..., D, D) -> (...)."""
        from .. import ops
        return ops.residual_eval(self.eval_params(jacobian.shape[-1]), jacobian.detach(), want_momentum=False,
                                 want_div=True)[1]

    def forward(self, *args) -> Tensor:
        """mse of the divergence against zero: a 0-d tensor."""
        from .. import ops
        return ops.mean_squares(self.func(*args).reshape(-1, 1))[0]


class _Momentum(_ResidualSpec):
    def func(self, internal_input, u: Tensor, u_jac: Tensor, u_laplace: Tensor, p_grad: Tensor) -> Tensor:
        """Momentum residual (..., D) from the internal-domain inputs (FoamData with cellToRegion [, d, f]), the
        velocity, its Jacobian and Laplacian terms, and the pressure gradient."""
        from .. import ops
        dims = u.shape[-1]
        zone = internal_input['cellToRegion']
        dcoef = fcoef = None
        if self.kind == 'variable':
            dcoef, fcoef = internal_input['d'], internal_input['f']
        elif self.kind == 'manufactured':
            fcoef = internal_input['f']
        det = lambda t: None if t is None else t.detach()
        return ops.residual_eval(self.eval_params(dims), det(u_jac), det(u), det(u_laplace), det(p_grad), det(zone),
                                 det(dcoef), det(fcoef))[0]

    def forward(self, *args) -> Tensor:
        """Per-component mean of the squared residual: (D,)."""
        from .. import ops
        return ops.mean_squares(self.func(*args))


class ContinuityLoss(_Continuity):
    """div U = 0 on raw outputs (reference models/losses.py:149-164)."""
    kind = 'manufactured'


class ContinuityLossStandardized(_Continuity):
    """div U = 0 on standardised outputs (reference models/losses.py:167-190)."""

    def __init__(self, u_scaler: StandardScaler, points_scaler: StandardScaler):
        super().__init__()
        self.u_scaler, self.points_scaler = u_scaler, points_scaler


class MomentumLossManufactured(_Momentum):
    """Raw-output momentum residual with forcing term (reference models/losses.py:193-225)."""
    kind = 'manufactured'

    def __init__(self, nu: float, d: float, f: float):
        super().__init__()
        self.nu, self.d, self.f = nu, d, f


class MomentumLossFixed(_Momentum):
    """Standardised outputs, scalar Darcy / Forchheimer coefficients (reference models/losses.py:228-270)."""
    kind = 'fixed'

    def __init__(self, nu: float, d: float, f: float, u_scaler: StandardScaler, points_scaler: StandardScaler,
                 p_scaler: StandardScaler):
        super().__init__()
        self.nu, self.d, self.f = nu, d, f
        self.u_scaler, self.points_scaler, self.p_scaler = u_scaler, points_scaler, p_scaler


class MomentumLossVariable(_Momentum):
    """Standardised outputs, per-point per-component coefficients (reference models/losses.py:273-319)."""
    kind = 'variable'

    def __init__(self, nu: float, u_scaler: StandardScaler, points_scaler: StandardScaler, p_scaler: StandardScaler,
                 d_scaler: Normalizer, f_scaler: Normalizer):
        super().__init__()
        self.nu = nu
        self.u_scaler, self.points_scaler, self.p_scaler = u_scaler, points_scaler, p_scaler
        self.d_scaler, self.f_scaler = d_scaler, f_scaler

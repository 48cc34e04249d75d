"""PorousPinnBase: the reference's training / validation / prediction seams
(reference models/model_base.py:69-254) on top of the CUDA executor.

`training_step(batch, batch_idx)` keeps its signature and returns a 0-d loss tensor whose
`.backward()` delivers the parameter gradients, but the body is one fused step
(engine.PinnExecutor.step): a forward-mode jet pass instead of the reference's 1 + D + D*D + 1
autograd sweeps, one residual kernel, one reverse pass.  The gradients are computed in the step
itself; `backward()` only scales them by the incoming gradient and hands them to autograd.

`laplacian` selects the viscous operator:
  'reference'  what training_step computes as written (`get_laplacian(points, U)`, :195)
  'true'       the documented operator (`get_laplacian(points, get_jacobian(points, U))`)
"""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import Tensor, nn

from .. import _lib
from .._lib import LAP_MODES, LOSS_KINDS, ResidualParams
from ..dataset.foam_data import FoamData
from ..engine import PinnExecutor
from .losses import LossLogger

try:  # a real LightningModule when lightning is installed, so Trainer.fit() accepts the model
    import lightning as _L
    _Base = _L.LightningModule
except Exception:  # lightning is not part of this image: same surface on top of nn.Module
    class _Base(nn.Module):
        def __init__(self):
            super().__init__()
            self.logged: dict = {}
            self.hparams_saved: dict = {}
            self.global_step = 0

        def save_hyperparameters(self, *args, **kwargs):
            pass

        def log(self, name, value, **kwargs):
            self.logged[name] = value


class _StepGradients(torch.autograd.Function):
    """Connects the fused step to autograd: forward returns the loss computed by the kernels,
    backward returns the gradients the same step already produced (scaled by grad_output)."""

    @staticmethod
    def forward(ctx, executor: PinnExecutor, loss: Tensor, *params):
        ctx.executor = executor
        return loss.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        ex = ctx.executor
        scaled = ex.flat_grad * grad_out
        ex.last_flat_grad = scaled      # every p.grad below is a view of this buffer: one all-reduce covers them all
        grads, off = [], 0
        for p in ex.params:
            grads.append(scaled[off:off + p.numel()].view_as(p))
            off += p.numel()
        return (None, None, *grads)


class _JetVJPPoints(torch.autograd.Function):
    """out[b,n,k] = sum_c g[b,n,c] * J1[b,n,c,k]: the reverse sweep `autograd.grad(y, points, g)` of a per-point model,
    read off the forward-mode jet.  Differentiable once more for the Hessian DIAGONAL (what get_laplacian consumes): the
    jet carries d2y_c/dx_k2 but no mixed partials, so entries of the second derivative that would need them are NaN
    instead of silently wrong.  `points` is an input only so that autograd routes that second derivative to it."""

    @staticmethod
    def forward(ctx, points, g, j1, j2):
        ctx.save_for_backward(g, j1, j2)
        return torch.einsum('bnc,bnck->bnk', g, j1)

    @staticmethod
    def backward(ctx, gg):
        g, j1, j2 = ctx.saved_tensors
        grad_g = torch.einsum('bnk,bnck->bnc', gg, j1)
        # d/dx_l of out = sum_c g_c sum_k gg_k d2y_c/(dx_k dx_l): the k == l term is known, any k != l with gg_k != 0
        # would need a mixed partial
        diag = torch.einsum('bnc,bnck->bnk', g, j2) * gg
        nz = (gg != 0).to(diag.dtype)
        mixed = nz.sum(-1, keepdim=True) - nz
        diag = torch.where(mixed > 0, torch.full_like(diag, float('nan')), diag)
        return diag, grad_g, None, None


class _JetForward(torch.autograd.Function):
    """Model.forward(autograd_points, x) as an autograd node over the points: forward runs ONE forward-mode jet pass
    (values, d/dx_k, d2/dx_k2) through the CUDA kernels, backward answers `autograd.grad(y, points, g)` from the stored
    jet instead of a reverse sweep through the network -- so the reference's helpers (calculate_gradients, get_jacobian,
    get_laplacian, models/model_base.py:11-53) and its `predict_step` derivative stack work on this model unchanged.
    Gradients with respect to the PARAMETERS do not flow through this node (training goes through training_step)."""

    @staticmethod
    def forward(ctx, points, model, x):
        ex = model.executor
        d = model.dims
        b, n, _ = points.shape
        yj = ex.forward_jets(points, x.data, x.labels, x.domain, order=2)
        y = yj.t[:, :, :d + 1].reshape(1 + 2 * d, b, n, d + 1)
        j1 = y[1:1 + d].permute(1, 2, 3, 0).contiguous()         # [b, n, c, k] = dy_c/dx_k
        j2 = y[1 + d:1 + 2 * d].permute(1, 2, 3, 0).contiguous()  # [b, n, c, k] = d2y_c/dx_k2
        ctx.save_for_backward(points, j1, j2)
        return y[0].clone()

    @staticmethod
    def backward(ctx, g):
        points, j1, j2 = ctx.saved_tensors
        return _JetVJPPoints.apply(points, g, j1, j2), None, None


def calculate_gradients(outputs: Tensor, inputs: Tensor) -> Tensor:
    """d(sum of outputs)/d(inputs), differentiable again (reference models/model_base.py:11-20)."""
    return torch.autograd.grad(outputs, inputs, grad_outputs=torch.ones_like(outputs), retain_graph=True,
                               create_graph=True)[0]


def get_jacobian(points: Tensor, u: Tensor) -> Tensor:
    """jac[..., i, j] = dU_i/dx_j, one sweep per velocity component (reference models/model_base.py:23-35)."""
    return torch.stack([calculate_gradients(u[..., i:i + 1], points) for i in range(points.shape[-1])], dim=-2)


def get_laplacian(points: Tensor, jacobian: Tensor) -> Tensor:
    """lap[..., i, j] = d/dx_j of jacobian[..., i, j] (reference models/model_base.py:38-53).  As in the reference,
    passing U instead of the Jacobian (what training_step does at :195) slices the POINT axis and yields first
    derivatives of the first D points."""
    dims = points.shape[-1]
    rows = []
    for i in range(dims):
        rows.append(torch.cat([calculate_gradients(jacobian[..., i:i + 1, j], points)[..., j:j + 1] for j in range(dims)], -1))
    return torch.stack(rows, dim=-2)


def enable_internal_autograd(batch: FoamData):
    """(internal points as an autograd leaf, all points) (reference models/model_base.py:56-66)."""
    internal_points = batch['internal']['C']
    internal_points.requires_grad = True
    return internal_points, torch.cat([internal_points, batch['boundary']['C']], dim=-2)


class PorousPinnBase(_Base):
    def __init__(self, out_features: int, enable_data_loss=True, loss_scaler=None, laplacian: str = 'reference'):
        super().__init__()
        self.verbose_predict = False
        self.cuda_graph = False        # training_step replays the fused step from a CUDA graph (per input signature)
        self.pipeline_geometry = False  # with cuda_graph: FPS / ball query of the announced next batch run beside this step
        self._next_batch = None
        self.enable_data_loss = bool(enable_data_loss)
        self.dims = out_features - 1
        self.laplacian = laplacian
        n = out_features
        physics = ['Continuity loss', 'Momentum x loss', 'Momentum y loss', 'Momentum z loss'][:n]
        boundary = ['Boundary loss p', 'Boundary loss ux', 'Boundary loss uy', 'Boundary loss uz'][:n]
        obs = ['Observations loss p', 'Observations loss ux', 'Observations loss uy',
               'Observations loss uz'][:n] if self.enable_data_loss else []
        errors = ['error p', 'error ux', 'error uy', 'error uz'][:n]
        self.training_loss_togger = LossLogger(self, 'Total loss', *physics, *boundary, *obs,
                                               *[f'Train {e}' for e in errors])
        self.val_loss_logger = LossLogger(self, *[f'Validation {e}' for e in errors])
        self.predicted_labels = self.get_predicted_labels()
        self.extra_labels = self.get_extra_labels()
        self.loss_scaler = loss_scaler
        self._executor: Optional[PinnExecutor] = None
        self._prm_cache: dict = {}
        self.last_step = None

    # ---- reference helpers ---------------------------------------------------------------------
    def to(self, *args: Any, **kwargs: Any):
        super().to(*args, **kwargs)
        if self.loss_scaler is not None:
            self.loss_scaler = self.loss_scaler.to(*args, **kwargs)
        self._executor = None
        self._prm_cache.clear()
        return self

    def get_predicted_labels(self) -> dict:
        u = ['Ux', 'Uy', 'Uz'][:self.dims]
        return {**dict.fromkeys(u), 'p': None, 'U': u}

    def get_extra_labels(self) -> dict:
        m = ['Momentumx', 'Momentumy', 'Momentumz'][:self.dims]
        return {**dict.fromkeys(m), 'div': None, 'Momentum': m}

    def postprocess_out(self, u: Tensor, p: Tensor):
        return u, p

    def transfer_batch_to_device(self, batch: FoamData, device, dataloader_idx: int = 0) -> FoamData:
        return batch.to(device, non_blocking=True)

    # ---- executor plumbing -----------------------------------------------------------------------
    @property
    def executor(self) -> PinnExecutor:
        if self._executor is None:
            self._executor = PinnExecutor(self)
        return self._executor

    def build_plan(self) -> dict:
        raise NotImplementedError

    def loss_spec(self) -> dict:
        """{'kind', 'nu', 'd', 'f', scalers...} -- provided by the concrete model."""
        raise NotImplementedError

    # attributes the residual parameters (and every captured step graph) are built from: assigning one of them drops the
    # cached parameter blocks and the captured graphs.  In-place edits of a scaler's tensors are not seen -- call
    # invalidate_residual_params() after such an edit.
    _PRM_ATTRS = frozenset({'loss_scaler', 'enable_data_loss', 'u_scaler', 'p_scaler', 'points_scaler', 'd_scaler',
                            'f_scaler', 'momentum_loss', 'continuity_loss', 'laplacian'})

    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name in PorousPinnBase._PRM_ATTRS and '_prm_cache' in self.__dict__:
            self.invalidate_residual_params()

    def invalidate_residual_params(self) -> None:
        self._prm_cache.clear()
        if self.__dict__.get('_executor') is not None:
            self._executor.reset_graphs()

    def residual_params(self, labels: dict, laplacian: str) -> ResidualParams:
        key = (tuple(labels), laplacian, bool(self.enable_data_loss))
        if key in self._prm_cache:
            return self._prm_cache[key]
        spec = self.loss_spec()
        d = self.dims
        keys = list(labels.keys())
        col = lambda name: keys.index(name)
        prm = ResidualParams()
        prm.dims, prm.loss_kind, prm.lap_mode = d, LOSS_KINDS[spec['kind']], LAP_MODES[laplacian]
        prm.enable_data_loss = 1 if self.enable_data_loss else 0
        prm.nu, prm.d, prm.f = float(spec['nu']), float(spec.get('d', 0.0)), float(spec.get('f', 0.0))

        def vec(x, n):
            t = torch.as_tensor(x, dtype=torch.float64).flatten().cpu()
            if t.numel() == 1:
                t = t.repeat(n)
            return [float(v) for v in t[:n]]

        ones, zeros = [1.0] * 3, [0.0] * 3
        if spec['kind'] != 'manufactured':
            prm.c_std[:d], prm.u_std[:d], prm.u_mean[:d] = vec(spec['C'].std, d), vec(spec['U'].std, d), vec(spec['U'].mean, d)
            prm.p_std, prm.p_mean = vec(spec['p'].std, 1)[0], vec(spec['p'].mean, 1)[0]
        else:
            prm.c_std[:], prm.u_std[:], prm.u_mean[:] = ones, ones, zeros
            prm.p_std, prm.p_mean = 1.0, 0.0
        if spec['kind'] == 'variable':
            prm.d_min[:d], prm.d_range[:d] = vec(spec['d_scaler'].min, d), vec(spec['d_scaler'].range, d)
            prm.f_min[:d], prm.f_range[:d] = vec(spec['f_scaler'].min, d), vec(spec['f_scaler'].range, d)
        u_names = labels['U']
        prm.col_u[:d] = [col(n) for n in u_names]
        prm.col_p, prm.col_zone = col('p'), col('cellToRegion')
        if spec['kind'] == 'variable':
            prm.col_d[:d] = [col(n) for n in labels['d']]
        if spec['kind'] in ('variable', 'manufactured'):
            prm.col_f[:d] = [col(n) for n in labels['f']]
        n_terms = 2 * d + 2 + ((d + 1) if self.enable_data_loss else 0)
        w = self.loss_scaler.weight_list(n_terms) if self.loss_scaler is not None else [1.0] * n_terms
        prm.weights[:] = (w + [0.0] * 16)[:16]
        self._prm_cache[key] = prm
        return prm

    # ---- model API ----------------------------------------------------------------------------------
    def forward(self, autograd_points: Tensor, x: FoamData) -> FoamData:
        """Predictions at `autograd_points` (B, N, D): FoamData with columns [U..., p] and x's domain."""
        if autograd_points.requires_grad and torch.is_grad_enabled():
            # derivative queries (get_jacobian / get_laplacian / calculate_gradients) are answered from the jet
            y = _JetForward.apply(autograd_points, self, x)
        else:
            y = self.executor.forward_values(autograd_points, x.data, x.labels, x.domain)
        return FoamData(y, self.predicted_labels, x.domain)

    def fused_step(self, batch: FoamData, laplacian: Optional[str] = None, keep_outputs: bool = False, geo=None,
                   accumulate: bool = False):
        """The hot path without the autograd wrapper: fills `executor.flat_grad`, returns StepResult.  `geo`: the
        batch's set-abstraction geometry when it was computed ahead (`executor.geometry`); `accumulate`: add this
        batch's gradient to the buffer (micro-batches of one optimizer step)."""
        if geo is None:
            geo = getattr(batch, 'geometry', None)       # a DeviceFoamDataset batch brings its cached geometry
        return self.executor.step(batch.data, batch.labels, batch.domain, laplacian or self.laplacian, keep_outputs, geo,
                                  accumulate)

    def announce_next_batch(self, batch: Optional[FoamData]) -> None:
        """Optional hint for `pipeline_geometry`: the batch the NEXT training_step call will receive (already on the
        device).  Its set-abstraction geometry is then computed inside this step's graph, off the critical path."""
        self._next_batch = batch

    def training_step(self, batch: FoamData, batch_idx: int = 0):
        if self.cuda_graph:
            nxt, self._next_batch = self._next_batch, None
            res = self.executor.graphed_step(batch.data, batch.labels, batch.domain, self.laplacian, next_batch=nxt,
                                             geo=getattr(batch, 'geometry', None))
        else:
            res = self.fused_step(batch)
        self.last_step = res
        loss = _StepGradients.apply(self.executor, res.loss, *self.executor.params)
        d = self.dims
        out = res.out
        self.training_loss_togger.log(len(batch.data), out[32], *out[16:16 + res.n_terms], out[36], *out[33:33 + d])
        return loss

    def validation_step(self, batch: FoamData, batch_idx: int = 0):
        predicted = self.forward(batch['C'], batch)
        u_error, p_error = self.calculate_errors(batch, predicted)
        self.val_loss_logger.log(len(batch.data), p_error, *u_error)

    def calculate_errors(self, target: FoamData, predicted: FoamData):
        """MAE of de-standardised U (D,) and p (logging only; reference models/model_base.py:168-180)."""
        pu, pp = self.postprocess_out(predicted['U'], predicted['p'])
        tu, tp = self.postprocess_out(target['U'], target['p'])
        return (pu - tu).abs().reshape(-1, pu.shape[-1]).mean(0), (pp - tp).abs().mean()

    def predict_step(self, batch: FoamData, batch_idx: int = 0):
        if self.verbose_predict:
            # (predicted, residuals): residuals = cat([momentum_error, div]) on the internal points
            # (reference models/model_base.py:233-252), from one jet forward + pcfd_residual_fields
            pred, fields = self.executor.predict_with_residuals(batch.data, batch.labels, batch.domain, self.laplacian)
            return (FoamData(pred, self.predicted_labels, batch.domain),
                    FoamData(fields, self.extra_labels, batch.domain))
        return self.forward(batch['C'], batch)

    def jets(self, batch: FoamData, laplacian: str = 'true') -> dict:
        """U, p, jacobian, laplacian and grad p at the internal points from the forward-mode jet
        (what get_jacobian / get_laplacian / calculate_gradients return in the reference)."""
        res = self.fused_step(batch, laplacian, keep_outputs=True)
        d = self.dims
        y = res.y_int.t[:, :, :d + 1]
        b = batch.data.shape[0]
        ni = y.shape[1] // b
        u, p = y[0, :, :d], y[0, :, d:]
        jac = y[1:1 + d, :, :d].permute(1, 2, 0)          # [row, i, j] = dU_i/dx_j
        dp = y[1:1 + d, :, d].permute(1, 0)
        out = {'U': u.reshape(b, ni, d), 'p': p.reshape(b, ni, 1), 'jacobian': jac.reshape(b, ni, d, d),
               'd_p': dp.reshape(b, ni, d)}
        if laplacian == 'true':
            out['laplacian'] = y[1 + d:1 + 2 * d, :, :d].permute(1, 2, 0).reshape(b, ni, d, d)
        return out

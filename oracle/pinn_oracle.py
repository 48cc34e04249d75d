"""TEST INFRASTRUCTURE ONLY -- stand-alone CPU restatement of the reference's PINN training step.

This is the oracle the CUDA path is checked against on the GPU box, where /root/reference does
not exist.  It restates, with plain torch CPU ops and reverse-mode autograd exactly as the
reference does, the algorithm of

  models/model_base.py:182-218   PorousPinnBase.training_step
  models/model_base.py:11-66     calculate_gradients / get_jacobian / get_laplacian /
                                 enable_internal_autograd
  models/losses.py:10-20,149-319 vector_loss, Continuity*, Momentum{Manufactured,Fixed,Variable}
  models/losses.py:39-61         FixedLossScaler
  models/modules.py:23-53,56-82,171-274   MLP, PointNetFeatureExtract, Branch, GeometryEncoder,
                                 NeuralOperator[Sequential]
  models/modules.py:94-139,277-325,403-423,483-527  the ++ set-abstraction stack (on top of the
                                 third-party restatements in oracle/pyg_restate.py)
  models/pipn/pipn_foam.py:59-166, models/pipn/pipn_baseline.py:12-124,
  models/pi_gano/{base,pi_gano,pi_gano_pp}.py          the six in-scope model forwards
  dataset/foam_data.py:36-61     FoamData label / sub-domain indexing

It is written as functions over a flat `{state_dict key: tensor}` mapping (the reference's own
parameter names) and a plain `spec` dict, not as a copy of the reference
classes.  It is PINNED: tests/golden/make_golden.py runs the unmodified reference (through
oracle/ref_shim.py) and this file on the same seeded inputs in the build container, asserts
agreement and commits the reference's outputs as fixtures; tests/test_oracle_golden.py re-checks
this file against those fixtures everywhere.  The ++ models are pinned only up to the third-party
ops (see oracle/pyg_restate.py: "parity unpinned" for fps / radius / PointNetConv themselves).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; it is never the thing measured as the product and never shipped.
"""
from __future__ import annotations

from typing import Callable

import torch
from torch import Tensor
from torch.nn import functional as F

from oracle import pyg_restate as pr

ACTS: dict[str, Callable[[Tensor], Tensor]] = {'silu': F.silu, 'tanh': torch.tanh}


# --------------------------------------------------------------------------------------
# FoamData indexing (dataset/foam_data.py:36-61) on a plain (data, labels, domain) triple
# --------------------------------------------------------------------------------------

def field(data: Tensor, labels: dict, name: str) -> Tensor:
    """Columns of a labelled tensor: a single label is one column (its position among the dict
    keys), a multi label is the concatenation of its sub-labels."""
    sub = labels[name]
    if sub:
        return torch.cat([field(data, labels, s) for s in sub], dim=-1)
    col = list(labels.keys()).index(name)
    return data[..., col:col + 1]


def rows(data: Tensor, ids: Tensor) -> Tensor:
    """Sub-domain rows of a batched (B, N, F) tensor: torch.gather with the ids repeated over F."""
    return torch.gather(data, 1, ids.unsqueeze(-1).repeat(1, 1, data.shape[-1]))


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------

def mlp(x: Tensor, p: dict, prefix: str, n_layers: int, act, last_activation: bool = True,
        dropout=None, training: bool = False, fmt: str = 'Linear {}', dropout_fn=None, site0: int = 0) -> Tensor:
    """models/modules.py:23-53: Linear -> act (-> Dropout) per layer; keys '<prefix>Linear k.*'.
    `dropout_fn(x, p, site)` (oracle/dropout_hash.CounterDropout) replaces torch's Philox mask by the CUDA path's
    counter-hash mask in training mode; layer k of this stack is layer site0 + k of the model's point chain, its
    dropped activations are consumed by layer site0 + k + 1."""
    for k in range(n_layers):
        name = prefix + fmt.format(k)
        x = F.linear(x, p[name + '.weight'], p[name + '.bias'])
        if k < n_layers - 1 or last_activation:
            x = act(x)
        if dropout is not None and dropout[k] > 0:
            if training and dropout_fn is not None:
                x = dropout_fn(x, dropout[k], site0 + k + 1)
            else:
                x = F.dropout(x, dropout[k], training)
    return x


def max_over_points(h: Tensor) -> Tensor:
    """torch.max(h, dim=1, keepdim=True)[0] (models/modules.py:81, :190, :214) with the margin probe."""
    if pr.MARGINS is not None:
        nb, n, c = h.shape
        pr.record_margin(h.reshape(nb * n, c), torch.arange(nb).repeat_interleave(n), nb)
    return torch.max(h, dim=1, keepdim=True)[0]


def set_abstraction_stack(x: Tensor, pos: Tensor, p: dict, prefix: str, spec: dict, act, tap=None) -> Tensor:
    """BatchedDecorator(SetAbstractionSeq) (models/modules.py:94-98, 483-527): flatten the batch,
    run every SetAbstraction (fps -> radius -> PointConvNext, :313-325, message :286-292) and the
    optional GlobalSetAbstraction (:412-423), return (B, n_out, E)."""
    nb, n = x.shape[0], x.shape[1]
    batch = torch.arange(nb).repeat_interleave(n)
    x, pos = x.reshape(nb * n, -1), pos.reshape(nb * n, -1)
    radii, fractions, channels = spec['radius'], spec['fraction'], spec['layers']
    k_max = spec['max_neighbors']
    for i, (frac, r, ch) in enumerate(zip(fractions, radii, channels)):
        idx = pr.fps(pos, batch, ratio=frac)
        row, col = pr.radius(pos, pos[idx], r, batch, batch[idx], max_num_neighbors=k_max)
        # edge_index = [col, row]; PointNetConv self-loop rule on the flattened bipartite graph
        keep = col != row
        src, dst = col[keep], row[keep]
        loops = torch.arange(min(pos.shape[0], idx.shape[0]))
        src, dst = torch.cat([src, loops]), torch.cat([dst, loops])
        pos_c = pos[idx]
        msg = torch.cat([x[src], pos[src] - pos_c[dst] / r], dim=1)   # operator precedence is the reference's
        msg = mlp(msg, p, f'{prefix}layers.Sa-{i}.conv.local_nn.', len(ch) - 1, act, True, fmt='lins.{}')
        x = pr.segment_max_first(msg, dst, idx.shape[0])
        if tap is not None:      # test hook: expose the per-level features (and their gradients)
            x.retain_grad()
            tap.append(x)
        pos, batch = pos_c, batch[idx]
    if len(channels) > len(radii):
        ch = channels[-1]
        h = mlp(torch.cat([x, pos], dim=1), p, f'{prefix}layers.Global-Sa.nn.', len(ch) - 1, act, True,
                fmt='lins.{}')
        x = pr.global_max_pool(h, batch, nb)
        batch = torch.arange(nb)
    return torch.stack(pr.unbatch(x, batch))


# --------------------------------------------------------------------------------------
# model forwards: (autograd_points (B,N,D), data (B,N,F)) -> y (B,N,D+1)
# --------------------------------------------------------------------------------------

def forward(spec: dict, p: dict, pts: Tensor, data: Tensor, labels: dict, domain: dict,
            training: bool = False, dropout_fn=None) -> Tensor:
    kind = spec['kind']
    act = ACTS[spec['activation']]
    n = pts.shape[-2]
    if kind in ('PipnFoam', 'PipnManufactured'):
        # models/pipn/pipn_foam.py:87-100, pipn_baseline.py:45-58; feature extractor :71-82.
        # PipnManufactured builds PointNetFeatureExtract with its default Tanh (pipn_baseline.py:39).
        fe_act = ACTS[spec.get('fe_activation', spec['activation'])]
        global_in = torch.cat([field(data, labels, 'boundaryId'), field(data, labels, 'sdf')], dim=-1)
        local = mlp(pts, p, 'feature_extract.local_feature.', len(spec['fe_local_layers']) - 1, fe_act)
        g = mlp(torch.cat([local, global_in], dim=-1), p, 'feature_extract.global_feature.',
                len(spec['fe_global_layers']) - 1, fe_act)
        g = max_over_points(g)
        seg_in = torch.cat([local, g.repeat(1, n, 1)], dim=-1)
        return mlp(seg_in, p, 'decoder.', len(spec['seg_layers']) - 1, act, False,
                   spec.get('seg_dropout'), training, dropout_fn=dropout_fn, site0=len(spec['fe_local_layers']) - 1)
    if kind in ('PipnFoamPp', 'PipnManufacturedPorousPp'):
        # models/pipn/pipn_foam.py:148-161, pipn_baseline.py:104-119 (feature order differs!)
        bnd = rows(data, domain['boundary'])
        bnd_c, bnd_id = field(bnd, labels, 'C'), field(bnd, labels, 'boundaryId')
        geom = torch.cat([bnd_c, bnd_id] if kind == 'PipnFoamPp' else [bnd_id, bnd_c], dim=-1)
        local = mlp(pts, p, 'feature_extract.local_feature.', len(spec['fe_local_layers']) - 1, act)
        g = set_abstraction_stack(geom, bnd_c, p, 'feature_extract.global_feature.module.',
                                  {'radius': spec['fe_radius'], 'fraction': spec['fe_fraction'],
                                   'layers': spec['fe_global_layers'],
                                   'max_neighbors': spec.get('max_neighbors', 64)}, act)
        seg_in = torch.cat([local, g.repeat(1, n, 1)], dim=-1)
        return mlp(seg_in, p, 'decoder.', len(spec['seg_layers']) - 1, act, False,
                   spec.get('seg_dropout'), training, dropout_fn=dropout_fn, site0=len(spec['fe_local_layers']) - 1)
    if kind in ('PiGano', 'PiGanoPp'):
        # models/pi_gano/pi_gano.py:49-69, pi_gano_pp.py:62-82, base.py:60-73
        par = []
        for sub in spec['variable_boundaries']['Subdomains']:
            sd = rows(data, domain[sub])
            par.append(torch.cat([field(sd, labels, 'C')] +
                                 [field(sd, labels, f) for f in spec['variable_boundaries']['Features']], dim=-1))
        par = torch.cat(par, dim=-2)
        if kind == 'PiGano':
            geom_in = torch.cat([field(data, labels, 'boundaryId'), field(data, labels, 'sdf'), pts.detach()], dim=-1)
            ge = mlp(geom_in, p, 'geometry_encoder.linear.', len(spec['geometry_layers']) - 1, act)
            ge = max_over_points(ge)
        else:
            bnd = rows(data, domain['boundary'])
            bnd_c = field(bnd, labels, 'C').detach()
            geom_in = torch.cat([bnd_c, field(bnd, labels, 'boundaryId')], dim=-1).detach()
            ge = set_abstraction_stack(geom_in, bnd_c, p, 'geometry_encoder.set_abstraction.module.',
                                       {'radius': spec['geometry_radius'], 'fraction': spec['geometry_fraction'],
                                        'layers': spec['geometry_layers'],
                                        'max_neighbors': spec.get('max_neighbors', 64)}, act)
        local = mlp(pts, p, 'points_encoder.', len(spec['local_layers']) - 1, act)
        h = torch.cat([local, ge.repeat(1, n, 1)], dim=-1)
        pe = mlp(par, p, 'branch.linear.', len(spec['branch_layers']) - 1, act)
        pe = max_over_points(pe)
        for k in range(spec['n_operators']):
            # models/modules.py:239-245: Dropout(act(Linear(h))) * par_embedding
            h = act(F.linear(h, p[f'neural_ops.Operator {k}.linear.0.weight'],
                             p[f'neural_ops.Operator {k}.linear.0.bias']))
            if spec['operator_dropout'][k] > 0:
                if training and dropout_fn is not None:     # operator k is layer (n_encoder + k) of the point chain
                    h = dropout_fn(h, spec['operator_dropout'][k], len(spec['local_layers']) - 1 + k + 1)
                else:
                    h = F.dropout(h, spec['operator_dropout'][k], training)
            h = h * pe
        return F.linear(h, p['reduction.weight'], p['reduction.bias'])
    raise KeyError(kind)


# --------------------------------------------------------------------------------------
# derivatives (models/model_base.py:11-53)
# --------------------------------------------------------------------------------------

def ones_vjp(out: Tensor, inp: Tensor) -> Tensor:
    return torch.autograd.grad(out, inp, grad_outputs=torch.ones_like(out), retain_graph=True,
                               create_graph=True)[0]


def jacobian_of(pts: Tensor, u: Tensor) -> Tensor:
    """jac[..., i, j] = d(sum U_i)/dx_j, one reverse sweep per output component."""
    return torch.stack([ones_vjp(u[..., d:d + 1], pts) for d in range(pts.shape[-1])], dim=-2)


def laplacian_of(pts: Tensor, second_arg: Tensor) -> Tensor:
    """get_laplacian exactly as written: D*D sweeps over `second_arg[..., i:i+1, j]`.

    Given the Jacobian (B,NI,D,D) this is lap[..., i, j] = d2 U_i / dx_j2 (the documented use).
    training_step passes U (B,NI,D) instead (models/model_base.py:195), in which case the slice
    addresses POINT i and the result is d U_j(point i) / d x_j(point n) -- reproduced verbatim by
    calling this function with U.
    """
    dims = pts.shape[-1]
    out = []
    for i in range(dims):
        comps = [ones_vjp(second_arg[..., i:i + 1, j], pts)[..., j:j + 1] for j in range(dims)]
        out.append(torch.cat(comps, -1))
    return torch.stack(out, dim=-2)


# --------------------------------------------------------------------------------------
# losses (models/losses.py)
# --------------------------------------------------------------------------------------

def per_component_mean(err: Tensor) -> Tensor:
    return err.reshape(-1, err.shape[-1]).mean(dim=0)


def continuity_residual(spec: dict, jac: Tensor) -> Tensor:
    diag = torch.diagonal(jac, 0, -1, -2)
    if spec['loss'] != 'manufactured':
        diag = diag * spec['scalers']['U_std'] / spec['scalers']['C_std']
    return diag.sum(-1)


def momentum_residual(spec: dict, internal: Tensor, labels: dict, u: Tensor, jac: Tensor, lap: Tensor,
                      dp: Tensor) -> Tensor:
    nu = spec['nu']
    zone = field(internal, labels, 'cellToRegion')
    if spec['loss'] == 'manufactured':
        # models/losses.py:209-217
        source = u * (spec['d'] * nu + 0.5 * torch.norm(u, dim=-1, keepdim=True) * spec['f'])
        conv = torch.matmul(jac, u.unsqueeze(-1)).squeeze(-1)
        visc = nu * lap.sum(-1)
        return conv - visc + dp + source * zone - field(internal, labels, 'f')
    sc = spec['scalers']
    u_raw = sc['U_std'] * u + sc['U_mean']
    if spec['loss'] == 'fixed':          # models/losses.py:256-266
        d_raw, f_raw = spec['d'], spec['f']
    else:                                # models/losses.py:301-311
        d_raw = sc['d_min'] + (sc['d_max'] - sc['d_min']) * field(internal, labels, 'd')
        f_raw = sc['f_min'] + (sc['f_max'] - sc['f_min']) * field(internal, labels, 'f')
    source = u_raw * (d_raw * nu + 0.5 * torch.norm(u_raw, dim=-1, keepdim=True) * f_raw)
    conv = torch.matmul(jac, (u_raw / sc['C_std']).unsqueeze(-1)).squeeze(-1) * sc['U_std']
    visc = nu * torch.matmul(lap, (1 / sc['C_std'] ** 2).unsqueeze(-1)).squeeze(-1) * sc['U_std']
    pres = (sc['p_std'] / sc['C_std']) * dp
    return conv - visc + pres + source * zone


def training_step(spec: dict, p: dict, data: Tensor, labels: dict, domain: dict,
                  laplacian: str = 'reference', training: bool = False, dropout_fn=None) -> dict:
    """One training step as models/model_base.py:182-218 performs it.

    laplacian='reference' reproduces the call as written (`get_laplacian(points, U)`);
    laplacian='true' is the documented operator (`get_laplacian(points, get_jacobian(points, U))`).
    Returns the scaled loss vector, its sum, the unscaled vector and the MAE log values.
    """
    dims = spec['dims']
    internal = rows(data, domain['internal'])
    boundary = rows(data, domain['boundary'])
    pts = field(internal, labels, 'C').detach().clone().requires_grad_(True)
    all_pts = torch.cat([pts, field(boundary, labels, 'C')], dim=-2)
    y = forward(spec, p, all_pts, data, labels, domain, training, dropout_fn)
    out_labels = {**{n: None for n in ['Ux', 'Uy', 'Uz'][:dims]}, 'p': None, 'U': ['Ux', 'Uy', 'Uz'][:dims]}

    y_int, y_bnd = rows(y, domain['internal']), rows(y, domain['boundary'])
    u_int, p_int = field(y_int, out_labels, 'U'), field(y_int, out_labels, 'p')
    bnd_p = F.mse_loss(field(y_bnd, out_labels, 'p'), field(boundary, labels, 'p'))
    bnd_u = per_component_mean((field(y_bnd, out_labels, 'U') - field(boundary, labels, 'U')) ** 2)

    jac = jacobian_of(pts, u_int)
    lap = laplacian_of(pts, u_int if laplacian == 'reference' else jac)
    dp = ones_vjp(p_int, pts)

    div = continuity_residual(spec, jac)
    cont = (div ** 2).mean()
    mom_field = momentum_residual(spec, internal, labels, u_int, jac, lap, dp)
    mom = per_component_mean(mom_field ** 2)

    terms = [cont, *mom, *bnd_u, bnd_p]
    if spec['enable_data_loss']:
        y_obs, t_obs = rows(y, domain['obs']), rows(data, domain['obs'])
        obs_u = per_component_mean((field(y_obs, out_labels, 'U') - field(t_obs, labels, 'U')) ** 2)
        obs_p = F.mse_loss(field(y_obs, out_labels, 'p'), field(t_obs, labels, 'p'))
        terms += [*obs_u, obs_p]
    unscaled = torch.stack(terms)
    w = spec.get('loss_weights')
    scaled = unscaled * torch.tensor(w, dtype=unscaled.dtype) if w is not None else unscaled
    loss = scaled.sum()

    # calculate_errors (models/model_base.py:168-180): MAE on inverse-transformed fields, all points
    u_all, p_all = field(y, out_labels, 'U'), field(y, out_labels, 'p')
    ut, pt = field(data, labels, 'U'), field(data, labels, 'p')
    if spec['loss'] != 'manufactured':
        sc = spec['scalers']
        u_all, ut = sc['U_std'] * u_all + sc['U_mean'], sc['U_std'] * ut + sc['U_mean']
        p_all, pt = sc['p_std'] * p_all + sc['p_mean'], sc['p_std'] * pt + sc['p_mean']
    u_err = per_component_mean((u_all - ut).abs())
    p_err = (p_all - pt).abs().mean()
    return {'loss': loss, 'losses': scaled, 'unscaled': unscaled, 'u_error': u_err.detach(),
            'p_error': p_err.detach(), 'y': y, 'jac': jac, 'lap': lap, 'dp': dp, 'points': pts,
            # predict_step(verbose) residual map (models/model_base.py:250): cat([momentum_error, div])
            'residuals': torch.cat([mom_field, div.reshape(*mom_field.shape[:-1], 1)], dim=-1).detach()}


def step_with_grads(spec: dict, p: dict, data: Tensor, labels: dict, domain: dict,
                    laplacian: str = 'reference', training: bool = False, dropout_fn=None) -> dict:
    """training_step + loss.backward(): returns the step outputs and {key: grad} for every parameter."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    out = training_step(spec, leaves, data, labels, domain, laplacian, training, dropout_fn)
    out['loss'].backward()
    out['grads'] = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return out


class Relobralo:
    """Restatement of RelobraloScaler.forward (reference models/losses.py:93-124) for a fixed Bernoulli draw
    `rho` (the reference samples torch.bernoulli(beta); beta = 1 -> rho = 1, beta = 0 -> rho = 0).
    Call once per training step with the unscaled loss vector; returns the weighted vector."""

    def __init__(self, num_losses: int, alpha=0.95, rho=1.0, tau=1.0, eps=1e-8, batch_size=1):
        self.n, self.alpha, self.rho, self.tau, self.eps, self.batch_size = num_losses, alpha, rho, tau, eps, batch_size
        self.init = torch.zeros(num_losses)
        self.prev = torch.zeros(num_losses)
        self.lam = torch.ones(num_losses)
        self.step = 0

    def __call__(self, losses: Tensor) -> Tensor:
        losses = losses.detach().float()
        step, self.step = self.step, self.step + 1
        if step == 0:                                             # :99-102
            self.init, self.prev = losses.clone(), losses.clone()
            return losses
        if step % self.batch_size == 0:                           # :106-119
            self.prev = self.prev / self.batch_size
            n_prev = (losses / (self.tau * self.prev)).max()
            n_init = (losses / (self.tau * self.init)).max()
            l_prev = torch.exp(losses / (self.tau * self.prev + self.eps) - n_prev)
            l_init = torch.exp(losses / (self.tau * self.init + self.eps) - n_init)
            l_prev = l_prev * (self.n / (l_prev.sum() + self.eps))
            l_init = l_init * (self.n / (l_init.sum() + self.eps))
            self.lam = self.alpha * (self.rho * self.lam + (1.0 - self.rho) * l_init) + (1.0 - self.alpha) * l_prev
            self.prev = losses.clone()
        else:                                                     # :126
            self.prev = self.prev + losses
        return self.lam * losses                                  # :127

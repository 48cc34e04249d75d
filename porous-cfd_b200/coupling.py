"""Max-pool coupling of vanilla PIPN (`PipnFoam`, `PipnManufactured`): the terms by which the reference's
"Jacobian" differs from the per-point Jacobian.

`calculate_gradients` (reference models/model_base.py:11-20) differentiates the SUM over all internal points of an
output component, and in vanilla PIPN every prediction depends on every point through the max-pooled global feature
g_c = max_n h_c(n) (models/modules.py:77-82).  So, with a(c) the arg-max row of feature c,

    jac_ref[n, i, k] = dU_i(n)/dx_k |_g  +  sum_{c: a(c) = n}  S_ic * H_ck(n)
    S_ic   = sum_{m internal} dU_i(m)/dg_c        (a reverse-mode quantity: D+1 value-only sweeps through the decoder)
    H_ck(n) = dh_c(n)/dx_k                        (forward jet of the local + global MLP at the internal points)

and likewise for grad p (i = D).  As written, `training_step` also builds its "Laplacian" from single-point sweeps
(`get_laplacian(points, U)`, models/model_base.py:195): lap_ref[n, i, j] = dU_j(point i)/dx_j(point n), which picks up
sum_{c: a(c) = n} s^(i)_jc * H_cj(n) with s^(i)_j = dU_j(point i)/dg, i < D.

Everything here is sequencing of the C-ABI kernels (jet layers, their reverse passes, column sums); the gathers at
the <= 1024 arg-max rows per geometry and the products of the small [B, 1024] factors are torch index ops.

Backward.  With G = dL/d(jac_ref, grad p) (what the residual kernel returns on the tangent planes of the output jet),
    lambda_ic = sum_k G[a(c), i, k] H_ck(a(c))    ->  dL/dS_ic
    mu_ck     = sum_i G[a(c), i, k] S_ic          ->  dL/dH_ck(a(c)): a (sparse) cotangent jet for the global MLP
and dL/dtheta through S is the gradient of  Q = sum_i sum_m  d/de U_i(m; p_m + e v_i),  v_i = W_g lambda_i: the
derivative of the decoder along a per-geometry direction of its first pre-activation -- a forward tangent, so a
second jet pass (direction channels instead of spatial ones) followed by the ordinary reverse pass delivers it.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from .ops import Jet


def _act_d(act: Optional[str], z: Tensor):
    """f'(z), f''(z) of the pending activation of the global MLP's last layer (torch glue on [B*G] values)."""
    if act is None:
        return torch.ones_like(z), torch.zeros_like(z)
    if act == 'silu':
        s = torch.sigmoid(z)
        t = s * (1 - s)
        return s + z * t, t * (2 + z * (1 - 2 * s))
    if act == 'tanh':
        t = torch.tanh(z)
        d = 1 - t * t
        return d, -2 * t * d
    raise KeyError(act)


def _plane0(j: Jet) -> Jet:
    return Jet(j.t[0:1], j.width)


def _concat_index(layers) -> int:
    return next(i for i, L in enumerate(layers) if L.cvec_key == 'concat')


def decoder_vjp(ex, zs, rows_per_geom: int, gz: Jet, salt_base: int, n_geom: int, keep_rows: Optional[Tensor] = None):
    """Value-only reverse sweep through the decoder layers above the concat layer: gz [1][rows][n_out] cotangent of
    the outputs -> per-geometry column sums of the cotangent of the concat layer's pre-activation, [n_geom, n_concat]
    (and, with `keep_rows`, that cotangent at the given rows).  No parameter gradient is touched."""
    from .engine import _tin
    ctx, layers = ex.ctx, ex.plan['point_layers']
    ci = _concat_index(layers)
    for li in range(len(layers) - 1, ci, -1):
        L = layers[li]
        gz = ops.jet_linear_bwd_dx(gz, L.weight, L.col_lo, _plane0(zs[li]), _tin(ctx, L, None, salt_base + li), None,
                                   rows_per_geom, L.k, L.n)
    Lc = layers[ci]
    out = torch.zeros((n_geom, ops.round4(Lc.n)), dtype=torch.float32, device=gz.t.device)
    ctx.need_workspace(ops.dw_workspace_bytes(1, gz.rows, rows_per_geom, Lc.k, Lc.n))
    ops.jet_linear_bwd_dw(gz, _plane0(zs[ci]), None, None, 0, None, out, rows_per_geom, Lc.k, Lc.n, ctx.workspace)
    if keep_rows is not None:
        return out, gz.t[0].index_select(0, keep_rows)
    return out


class Coupling:
    """State of one step's coupling terms (forward results needed by the backward)."""

    def __init__(self):
        self.zs_h = None        # jets of the global MLP at the internal points
        self.rows = None        # [B*G] flattened internal row of the arg-max, 0 where masked
        self.cols = None        # [B*G] feature index
        self.mask = None        # [B*G] 1.0 where the arg-max is an internal point
        self.hz = None          # [cj, B*G] pre-activation jet of h at the arg-max rows
        self.S = None           # [D+1, B*G]
        self.gcv = None         # [D+1][B, n_concat] column sums of the decoder cotangents
        self.s1 = None          # [D(point i), D(output j), B*G] single-point sensitivities dU_j(point i)/dg_c
        self.gcv1 = None        # [D][D][B, n_concat]
        self.w = None           # [D] 1 / c_std^2 (weights of the Laplacian row sum)
        self.visc_extra = None  # [B*NI, D] handed to the residual kernel
        self.gvisc = None       # [B*NI, D] returned by it
        self.local_row = None   # [B*G] arg-max row within its geometry


def forward(ex, data: Tensor, labels: dict, int_ids: Tensor, zs_int, saved: dict, cj: int, c_std) -> Coupling:
    """Adds the coupling terms to the tangent planes of the output jet zs_int[-1] (in place)."""
    plan, ctx = ex.plan, ex.ctx
    b, n_rows, _ = data.shape
    ni = int_ids.shape[1]
    d = plan['dims']
    layers = plan['point_layers']
    nl = len(plan['local_layers'])
    glayers, gpend = plan['global_layers'], plan['global_pending_act']
    lw = plan['local_layers'][-1].n
    g_width = glayers[-1].n
    st = Coupling()

    # ---- H: jets of the global MLP at the internal points (input = [local jets | boundaryId, sdf (zero tangents)])
    gcols = ex._cols(labels, 'boundaryId') + ex._cols(labels, 'sdf')
    gin = Jet.empty(cj, b * ni, lw + len(gcols), data.device)
    ops.zero_(gin.t)
    zl = zs_int[nl]
    for c in range(cj):
        ops.gather_cols(zl.t[c], 1, b * ni, zl.ld, None, 0, b * ni, list(range(lw)), gin.t[c], gin.ld, b * ni)
    ex._gather(data, int_ids, ni, gcols, gin.t[0], gin.ld, ni, 0, lw)
    from .engine import chain_forward
    st.zs_h = chain_forward(ctx, glayers, gin, ni)

    arg = saved['pipn']['global']['arg'].long()[:, :g_width]                     # [B, G] row within the geometry (all N points)
    internal = arg < ni
    st.mask = internal.reshape(-1).float()
    base = (torch.arange(b, device=data.device) * ni)[:, None]
    st.rows = torch.where(internal, arg + base, base.expand_as(arg)).reshape(-1)
    st.cols = torch.arange(g_width, device=data.device).repeat(b)
    st.hz = st.zs_h[-1].t[:, st.rows, st.cols]                                   # [cj, B*G]

    # ---- S_i = sum over internal points of dU_i / dg: D+1 value-only sweeps through the decoder
    y = zs_int[-1]
    cl = plan['concat_layer']
    gjet = saved['gjet']
    S, st.gcv = [], []
    geom_first = torch.arange(b, device=data.device) * ni
    first_rows = (geom_first[None, :] + torch.arange(d, device=data.device)[:, None]).reshape(-1)   # (point i, geometry b)
    single = []       # per output j: cotangent of the concat pre-activation at the first D points of every geometry
    for i in range(d + 1):
        gz = Jet.empty(1, y.rows, y.width, data.device)
        ops.zero_(gz.t)
        gz.t[0, :, i] = 1.0
        gcv, r_first = decoder_vjp(ex, zs_int, ni, gz, 100, b, first_rows)
        if i < d:
            single.append(r_first)      # the decoder is per-point: row (b, i) of this sweep IS dU_i... /dp at that single point
        st.gcv.append(gcv)
        s_i = ops.jet_linear_bwd_dx(Jet(gcv.unsqueeze(0), cl.n), cl.weight, cl.col_lo, gjet, None, None, 0, cl.k, cl.n)
        S.append(s_i.t[0, :, :g_width].reshape(-1))
    st.S = torch.stack(S)                                                        # [D+1, B*G]

    # ---- jac_ref, grad p: add S_ic * H_ck at the arg-max rows
    f1, _ = _act_d(gpend, st.hz[0])
    hk = f1 * st.hz[1:1 + d] * st.mask                                           # [D, B*G]
    contrib = st.S[:, None, :] * hk[None, :, :]                                  # [i, k, B*G]
    ii = torch.arange(d + 1, device=data.device)[:, None, None].expand_as(contrib)
    kk = (1 + torch.arange(d, device=data.device))[None, :, None].expand_as(contrib)
    rr = st.rows[None, None, :].expand_as(contrib)
    _scatter_add(y.t, (y.t.stride(0), y.t.stride(1), 1), (kk, rr, ii), contrib)

    # ---- the Laplacian as written (get_laplacian(points, U), models/model_base.py:195): lap[n, i, j] =
    #      dU_j(point i)/dx_j(point n) = delta(n, i) dU_j(i)/dx_j |_g + sum_{c: a(c)=n} s^(i)_jc H_cj(n).
    #      The residual kernel forms the first term from the output jet at rows n < D, which now holds
    #      jac_ref = per-point + C, so C's diagonal is taken out again there.
    if getattr(ex, '_coupling_w', None) is None:     # host -> device once (not capturable in a CUDA graph)
        ex._coupling_w = torch.tensor([1.0 / float(v) ** 2 for v in c_std[:d]], dtype=torch.float32, device=data.device)
    st.w = ex._coupling_w
    st.local_row = torch.where(internal, arg, torch.zeros_like(arg)).reshape(-1)
    # s^(i)_j = dU_j(point i)/dg: the sweep for output j above already holds dU_j(m)/dp_m at EVERY row m (the decoder
    # acts per point), so the single-point sensitivities are its rows (b, i) times W_g -- no further sweeps
    ldc = ops.round4(cl.n)
    r_all = torch.stack(single)                                                  # [j, D*B (i major), ld]
    r_all = r_all.reshape(d, d, b, -1).permute(1, 0, 2, 3).contiguous()          # [i, j, B, ld]
    st.gcv1 = [[r_all[i, j] for j in range(d)] for i in range(d)]
    s_flat = ops.jet_linear_bwd_dx(Jet(r_all.reshape(1, d * d * b, -1), cl.n), cl.weight, cl.col_lo,
                                   Jet(torch.empty((1, d * d * b, ops.round4(cl.k)), dtype=torch.float32, device=data.device), cl.k),
                                   None, None, 0, cl.k, cl.n)
    s1 = s_flat.t[0, :, :g_width].reshape(d, d, b * g_width)
    st.s1 = s1                                                                   # [i, j, B*G]
    vx = torch.zeros((b * ni, d), dtype=torch.float32, device=data.device)
    add = (st.s1 * (st.w[None, :, None] * hk[None, :, :])).sum(1)               # [i, B*G]: sum_j w_j s^(i)_jc H_cj
    _scatter_add(vx, (d, 1), (st.rows[None, :].expand_as(add), torch.arange(d, device=data.device)[:, None].expand_as(add)), add)
    # minus the coupling part of the diagonal the kernel reads at rows n = i < D:  sum_j w_j C[n, j, j]
    diag = torch.stack([contrib[j, j] for j in range(d)])                         # [j, B*G] = C[a(c), j, j] per (b, c)
    corr = -(st.w[:, None] * diag).sum(0) * (st.local_row < d).float() * st.mask  # [B*G]
    _scatter_add(vx, (d, 1), (st.rows, st.local_row.clamp(max=d - 1)), corr)
    st.visc_extra = vx
    st.gvisc = torch.empty_like(vx)
    return st


def backward(ex, st: Coupling, data: Tensor, int_ids: Tensor, zs_int, gy_int: Jet, saved: dict, gcvecs: dict) -> None:
    """Parameter gradients through the coupling terms (module docstring).  Call after the residual kernel and before
    the encoder backward (it accumulates into gcvecs['concat'])."""
    from .engine import chain_backward, chain_forward
    plan, ctx = ex.plan, ex.ctx
    b = data.shape[0]
    ni = int_ids.shape[1]
    d = plan['dims']
    dev = data.device
    layers = plan['point_layers']
    ci = _concat_index(layers)
    nl = len(plan['local_layers'])
    glayers, gpend = plan['global_layers'], plan['global_pending_act']
    lw = plan['local_layers'][-1].n
    g_width = glayers[-1].n
    cl = plan['concat_layer']
    bg = st.rows.numel()

    # ---- cotangents at the arg-max rows
    f1, f2 = _act_d(gpend, st.hz[0])
    hk = f1 * st.hz[1:1 + d] * st.mask                                           # [k, BG]
    G = torch.stack([gy_int.t[1:1 + d, st.rows, i] for i in range(d + 1)])       # [i, k, BG] = dL/d jac_ref[a(c), i, k]
    gvx = st.gvisc[st.rows]                                                      # [BG, D(i)]
    # the diagonal correction  -[n == i] sum_j w_j C[n, j, j]  hands  -w_j gvisc[n, n]  to C[n, j, j]
    small = ((st.local_row < d).float() * st.mask)
    gdiag = gvx.gather(1, st.local_row.clamp(max=d - 1)[:, None])[:, 0] * small  # gvisc[a(c), local row] where local row < D
    for j in range(d):
        G[j, j] = G[j, j] - st.w[j] * gdiag
    lam = (G * hk[None]).sum(1)                                                  # [i, BG]   dL/dS_ic
    mu = (G * st.S[:, None, :]).sum(0)                                           # [k, BG]   dL/dH_ck
    lam1 = st.w[None, :, None] * gvx.t()[:, None, :] * hk[None, :, :]            # [i, j, BG] dL/ds^(i)_jc
    mu = mu + (st.w[None, :, None] * gvx.t()[:, None, :] * st.s1).sum(0)         # + sum_i w_j gvisc[a(c), i] s^(i)_jc
    mu = mu * st.mask

    # ---- S = gcv @ W_g:  dW_g += gcv^T lambda,  v = lambda @ W_g^T (direction of the decoder's first pre-activation)
    lam_all = torch.cat([lam.reshape(d + 1, b, g_width), lam1.reshape(d * d, b, g_width)]).reshape(-1, g_width).contiguous()
    gcv_all = torch.cat(list(st.gcv) + [st.gcv1[i][j] for i in range(d) for j in range(d)])     # [(D+1+D*D)*B, ld]
    lam_jet = Jet(_pad4(lam_all).unsqueeze(0), g_width)
    ctx.need_workspace(ops.dw_workspace_bytes(1, lam_jet.rows, 0, cl.k, cl.n))
    ops.jet_linear_bwd_dw(Jet(gcv_all.unsqueeze(0), cl.n), lam_jet, None, ctx.grad(cl.weight), cl.col_lo, None, None, 0,
                          cl.k, cl.n, ctx.workspace)
    v_all = ops.jet_linear_fwd(lam_jet, None, cl.weight, cl.col_lo, cl.k, None, None, 0, cl.n).t[0]   # [(..)*B, ld]
    n_c = cl.n
    v = v_all[:(d + 1) * b, :n_c].reshape(d + 1, b, n_c)                         # direction for output i, all rows of b
    v1 = v_all[(d + 1) * b:, :n_c].reshape(d, d, b, n_c)                         # [i, j]: direction for output j at row (b, i)

    # ---- Q = sum_i sum_m d/de U_i(m; p_m + e * seed_i(m)): tangent pass through the decoder above the concat layer
    p_jet = zs_int[ci + 1]
    upper = layers[ci + 1:]
    geom_first = torch.arange(b, device=dev) * ni
    gp0 = torch.zeros((p_jet.rows, p_jet.ld), dtype=torch.float32, device=dev)
    outs = list(range(d + 1))
    for lo in range(0, d + 1, 3):
        grp = outs[lo:lo + 3]
        cjq = 4 if len(grp) == 3 else 3
        zq = Jet.empty(cjq, p_jet.rows, p_jet.width, dev)
        ops.zero_(zq.t)
        zq.t[0].copy_(p_jet.t[0])
        for slot, o in enumerate(grp):
            zq.t[1 + slot, :, :n_c] = v[o].repeat_interleave(ni, dim=0)
            if o < d:
                for i in range(d):
                    zq.t[1 + slot, geom_first + i, :n_c] += v1[i, o]
        zsq = chain_forward(ctx, upper, zq, ni, None, None, salt_base=100 + ci + 1)
        gq = Jet.empty(cjq, zsq[-1].rows, zsq[-1].width, dev)
        ops.zero_(gq.t)
        for slot, o in enumerate(grp):
            gq.t[1 + slot, :, o] = 1.0
        gp = chain_backward(ctx, upper, zsq, gq, ni, None, None, None, need_input_grad=True, salt_base=100 + ci + 1)
        gp0 += gp.t[0]
    # below the concat layer Q depends on the parameters through the VALUE of p only
    lower = layers[:ci + 1]
    chain_backward(ctx, lower, [_plane0(z) for z in zs_int[:ci + 1]], Jet(gp0.unsqueeze(0), p_jet.width), ni, None, None,
                   gcvecs, salt_base=100)

    # ---- H = f'(z0) z_k of the global MLP: cotangent jet (non-zero at the arg-max rows only), reverse through the
    #      global and local MLP jets
    zh = st.zs_h[-1]
    gh = Jet.empty(zh.cj, zh.rows, zh.width, dev)
    ops.zero_(gh.t)
    g0 = (mu * f2 * st.hz[1:1 + d]).sum(0) * st.mask
    _scatter_add(gh.t, (gh.t.stride(0), gh.t.stride(1), 1), (torch.zeros(bg, dtype=torch.long, device=dev), st.rows, st.cols), g0)
    kk = (1 + torch.arange(d, device=dev))[:, None].expand(d, bg)
    _scatter_add(gh.t, (gh.t.stride(0), gh.t.stride(1), 1), (kk, st.rows[None].expand(d, bg), st.cols[None].expand(d, bg)), mu * f1)
    gin = chain_backward(ctx, glayers, st.zs_h, gh, ni, need_input_grad=True)
    chain_backward(ctx, plan['local_layers'], zs_int[:nl + 1], Jet(gin.t, lw), ni, salt_base=100)


def _scatter_add(target: Tensor, strides, index_parts, values: Tensor) -> None:
    """target[i0, i1, ...] += values at broadcast index tensors `index_parts`, as ONE atomic-add kernel on the flattened
    target (index_put_(accumulate=True) sorts the indices first: a radix sort, a bounds reduction and an assert launch per
    call).  `strides`: element strides of the contiguous target."""
    lin = None
    for st_, ix in zip(strides, index_parts):
        term = ix * st_
        lin = term if lin is None else lin + term
    target.view(-1).index_add_(0, lin.reshape(-1), values.reshape(-1))


def _pad4(t: Tensor) -> Tensor:
    w = t.shape[-1]
    if w % 4 == 0:
        return t.contiguous()
    out = torch.zeros((*t.shape[:-1], ops.round4(w)), dtype=t.dtype, device=t.device)
    out[..., :w] = t
    return out

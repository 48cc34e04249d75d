"""Op-level parity of the CUDA kernels, called through the C ABI, against the oracle (index ops:
bit-exact) and against an fp64 torch restatement of the same algebra (floating-point ops)."""
import math

import numpy as np
import pytest
import torch

from helpers import GOLDEN, max_rel, rel_l2
from oracle import pyg_restate

pytestmark = pytest.mark.gpu

TOL = 1e-4  # relative, fp32 kernels vs fp64 restatement (north-star tolerance)
GEMM_TOL = 3e-5  # per-layer budget inside that gate (3xTF32 tensor-core engine: ~1e-5 on long row reductions)


@pytest.fixture(scope='module')
def ops():
    from porous_cfd_b200 import ops as o
    return o


def dev(t):
    return t.cuda()


# ---------------------------------------------------------------------------------------------
# index ops: bit-exact
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_fps_and_ball_query_match_fixture_bit_exact(ops, tag):
    z = np.load(f'{GOLDEN}/index_ops.npz')
    nb, n, d, ratio, r, k = z[f'{tag}/meta']
    nb, n, d, k = int(nb), int(n), int(d), int(k)
    pos = torch.from_numpy(z[f'{tag}/pos']).reshape(nb, n, d)
    idx = ops.fps(dev(pos), float(ratio))
    assert torch.equal(idx.cpu().flatten(), torch.from_numpy(z[f'{tag}/fps']))
    nbr, count = ops.ball_query(dev(pos), idx, float(r), k)
    row, col = torch.from_numpy(z[f'{tag}/row']), torch.from_numpy(z[f'{tag}/col'])
    nbr = nbr.cpu()
    got_row, got_slot = torch.nonzero(nbr >= 0, as_tuple=True)
    assert torch.equal(got_row, row)
    assert torch.equal(nbr[got_row, got_slot].long(), col)
    assert torch.equal(count.cpu().long(), torch.bincount(row, minlength=nbr.shape[0]))


@pytest.mark.parametrize('shape', [(4, 1000, 3, 0.5), (3, 777, 2, 0.25), (2, 4096, 3, 0.25), (1, 8192, 3, 0.5),
                                   (5, 33, 3, 0.5), (2, 1, 3, 1.0), (2, 20000, 3, 0.02), (1, 30000, 2, 0.01)])
def test_fps_bit_exact_against_oracle(ops, shape):
    nb, n, d, ratio = shape
    g = torch.Generator().manual_seed(n)
    pos = torch.rand(nb, n, d, generator=g) * 2 - 1
    want = pyg_restate.fps(pos.reshape(-1, d), torch.arange(nb).repeat_interleave(n), ratio)
    got = ops.fps(dev(pos), ratio).cpu().flatten()
    assert torch.equal(got, want)
    # properties: m = ceil(ratio n) distinct samples per geometry, first one is point 0
    m = math.ceil(ratio * n)
    per = got.reshape(nb, m) - (torch.arange(nb) * n)[:, None]
    assert all(len(set(r.tolist())) == m for r in per) and bool((per[:, 0] == 0).all())


@pytest.mark.parametrize('shape', [(2, 4096, 3, 0.25), (1, 8192, 3, 0.5), (2, 20000, 2, 0.02)])
def test_fps_cluster_kernel_bit_exact(ops, shape):
    """The thread-block-cluster / distributed-shared-memory variant (opt-in, PCFD_FPS_CLUSTER=1) in a fresh process."""
    import os
    import subprocess
    import sys
    nb, n, d, ratio = shape
    code = f'''
import sys, torch
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import pcfd_import; pcfd_import.load()
from porous_cfd_b200 import ops
from oracle import pyg_restate
g = torch.Generator().manual_seed({n})
pos = torch.rand({nb}, {n}, {d}, generator=g) * 2 - 1
want = pyg_restate.fps(pos.reshape(-1, {d}), torch.arange({nb}).repeat_interleave({n}), {ratio})
got = ops.fps(pos.cuda(), {ratio}).cpu().flatten()
assert torch.equal(got, want)
print("cluster fps ok")
'''
    env = dict(os.environ, PCFD_FPS_CLUSTER='1')
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'cluster fps ok' in r.stdout, r.stderr[-2000:]


def _surface_points(nb, n, d, gen):
    """Points on a few thin shells / curves with exact duplicates and a grid-aligned patch: what boundary point sets look
    like (highly non-uniform in the bounding box), plus the tie / zero-distance cases of the bucketed sampler."""
    t = torch.rand(nb, n, generator=gen) * 6.2831853
    r = 0.3 + 0.5 * (torch.rand(nb, n, generator=gen) > 0.5).float()
    pos = torch.zeros(nb, n, d)
    pos[..., 0] = r * torch.cos(t)
    pos[..., 1] = r * torch.sin(t)
    if d == 3:
        pos[..., 2] = torch.round(torch.rand(nb, n, generator=gen) * 8) / 8 - 0.5
    q = n // 8
    pos[:, :q] = torch.round(pos[:, :q] * 16) / 16                      # lattice points: many exact ties
    pos[:, q:2 * q] = pos[:, :q]                                        # exact duplicates
    return pos


@pytest.mark.parametrize('shape', [(2, 4096, 3, 0.5), (3, 5000, 2, 0.25), (2, 16384, 3, 0.5), (1, 65536, 2, 0.5),
                                   (1, 40000, 3, 0.05)])
@pytest.mark.parametrize('kind', ['uniform', 'surface'])
def test_fps_bucketed_bit_exact(ops, shape, kind):
    """Large point sets take the bucketed sampler (fps_bucket.cu): same selection, in the same order, as the plain
    algorithm of the oracle -- including ties and duplicates (lowest ORIGINAL index wins although storage is permuted)."""
    nb, n, d, ratio = shape
    g = torch.Generator().manual_seed(n + d)
    pos = torch.rand(nb, n, d, generator=g) * 2 - 1 if kind == 'uniform' else _surface_points(nb, n, d, g)
    want = pyg_restate.fps(pos.reshape(-1, d), torch.arange(nb).repeat_interleave(n), ratio)
    got = ops.fps(dev(pos), ratio).cpu().flatten()
    assert torch.equal(got, want)


def test_fps_bucketed_equals_plain_kernels(ops):
    """The same inputs through the plain one-CTA kernels (PCFD_FPS_BUCKET_MIN=0, fresh process) and the bucketed path."""
    import os
    import subprocess
    import sys
    code = f'''
import sys, torch
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import pcfd_import; pcfd_import.load()
from porous_cfd_b200 import ops
g = torch.Generator().manual_seed(11)
pos = torch.rand(2, 8192, 3, generator=g) * 2 - 1
torch.save(ops.fps(pos.cuda(), 0.5).cpu(), sys.argv[1])
'''
    import tempfile
    outs = []
    for mn in ('0', '4096'):
        with tempfile.NamedTemporaryFile(suffix='.pt') as f:
            r = subprocess.run([sys.executable, '-c', code, f.name], env=dict(os.environ, PCFD_FPS_BUCKET_MIN=mn),
                               capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(torch.load(f.name))
    assert torch.equal(outs[0], outs[1])


def test_fps_ties_pick_lowest_index(ops):
    pos = torch.tensor([[[0., 0.], [1., 0.], [1., 0.], [0., 1.], [0., 1.], [0.5, 0.5]]])
    assert ops.fps(dev(pos), 0.5).cpu().flatten().tolist() == [0, 1, 3]
    # more samples than distinct points: zero-distance ties still resolve to the lowest unused... index 0 repeats
    pos = torch.zeros(1, 40, 3)
    pos[0, 20:] = 1.0
    want = pyg_restate.fps(pos.reshape(-1, 3), torch.zeros(40, dtype=torch.long), 0.25)
    assert torch.equal(ops.fps(dev(pos), 0.25).cpu().flatten(), want)


@pytest.mark.parametrize('shape', [(3, 500, 3, 0.5, 0.5, 16), (2, 1000, 2, 0.25, 0.3, 64), (2, 300, 3, 0.5, 3.0, 8)])
def test_ball_query_and_edges_against_oracle(ops, shape):
    nb, n, d, ratio, r, k = shape
    g = torch.Generator().manual_seed(7 * n)
    pos = torch.rand(nb, n, d, generator=g) * 2 - 1
    flat = pos.reshape(-1, d)
    batch = torch.arange(nb).repeat_interleave(n)
    idx = pyg_restate.fps(flat, batch, ratio)
    row, col = pyg_restate.radius(flat, flat[idx], r, batch, batch[idx], k)
    nbr, _ = ops.ball_query(dev(pos), dev(idx.reshape(nb, -1)), r, k)
    got_row, got_slot = torch.nonzero(nbr.cpu() >= 0, as_tuple=True)
    assert torch.equal(got_row, row) and torch.equal(nbr.cpu()[got_row, got_slot].long(), col)
    # PyG self-loop rule on the flattened bipartite graph
    slots = ops.sa_edges(nbr, nb * n).cpu()
    keep = col != row
    loops = torch.arange(idx.numel())
    src, dst = torch.cat([col[keep], loops]), torch.cat([row[keep], loops])
    want = sorted(zip(dst.tolist(), src.tolist()))
    r2, s2 = torch.nonzero(slots >= 0, as_tuple=True)
    got = sorted(zip(r2.tolist(), slots[r2, s2].tolist()))
    assert got == want


def test_sa_gather_matches_message_inputs(ops):
    nb, n, d, f = 2, 64, 3, 5
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(nb, n, d, generator=g) * 2 - 1
    x = torch.randn(nb * n, 8, generator=g)
    idx = ops.fps(dev(pos), 0.5)
    nbr, _ = ops.ball_query(dev(pos), idx, 0.7, 4)
    slots = ops.sa_edges(nbr, nb * n)
    ein = ops.sa_gather(dev(x), 8, f, dev(pos), idx, slots, 0.7).cpu()[0]
    s, flat, ci = slots.cpu().long(), pos.reshape(-1, d), idx.cpu().flatten()
    for e in range(s.numel()):
        i, j = e // s.shape[1], int(s.flatten()[e])
        if j < 0:
            assert float(ein[e].abs().max()) == 0.0
        else:
            want = torch.cat([x[j, :f], flat[j] - flat[ci[i]] / 0.7])
            assert torch.equal(ein[e, :f + d], want)


# ---------------------------------------------------------------------------------------------
# segmented max / arg-max
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize('act', ['silu', 'tanh', None])
@pytest.mark.parametrize('shape', [(6, 17, 40), (3, 2500, 100), (1, 1, 7)])
def test_segmax_forward_backward(ops, act, shape):
    n_seg, seg_len, c = shape
    g = torch.Generator().manual_seed(1)
    z = torch.randn(n_seg * seg_len, c, generator=g) * 2
    z[3 % (n_seg * seg_len)] = z[0]          # exact ties: lowest slot must win
    slots = (torch.rand(n_seg, seg_len, generator=g) > 0.2).int() - 1 if seg_len > 1 else None
    f = {'silu': torch.nn.functional.silu, 'tanh': torch.tanh, None: lambda v: v}[act]
    zr = z.double().requires_grad_(True)
    a = f(zr).reshape(n_seg, seg_len, c)
    if slots is not None:
        valid = (slots >= 0)[:, :, None]
        a = torch.where(valid, a, torch.full_like(a, -float('inf')))
    want, warg = torch.max(a, dim=1)
    empty = torch.isinf(want)
    want = torch.where(empty, torch.zeros_like(want), want)
    zd = dev(z)
    out, arg = ops.segmax_fwd(zd, act, dev(slots) if slots is not None else None, n_seg, seg_len, c)
    assert max_rel(out[:, :c], want) < 1e-6 or float((out[:, :c].cpu() - want).abs().max()) < 1e-6
    assert torch.equal(torch.where(empty, torch.full_like(warg, -1), warg), arg.cpu().long())
    gout = torch.randn(n_seg, c, generator=g)
    want.backward(gout.double() * (~empty))
    gz = ops.segmax_bwd(dev(gout), c, arg, zd, act, n_seg, seg_len, c)[0, :, :c].cpu()
    assert float((gz.double() - zr.grad).abs().max()) < 1e-5


# ---------------------------------------------------------------------------------------------
# jet linear layer: forward, backward-to-input, backward-to-weights
# ---------------------------------------------------------------------------------------------

def act_jet_ref(act, z, dims, order, s):
    """fp64 torch restatement of the input transform; differentiable."""
    if act is None:
        return z * s
    z0 = z[0]
    if act == 'silu':
        sg = torch.sigmoid(z0)
        f0, d1 = z0 * sg, sg + z0 * sg * (1 - sg)
        d2 = sg * (1 - sg) * (2 + z0 * (1 - 2 * sg))
    else:
        t = torch.tanh(z0)
        f0, d1 = t, 1 - t * t
        d2 = -2 * t * (1 - t * t)
    # the closed forms above are checked against autograd of the activation itself
    zz = z0.detach().clone().requires_grad_(True)
    a0 = torch.nn.functional.silu(zz) if act == 'silu' else torch.tanh(zz)
    a1 = torch.autograd.grad(a0.sum(), zz, create_graph=True)[0]
    a2 = torch.autograd.grad(a1.sum(), zz)[0]
    assert torch.allclose(d1, a1.detach(), atol=1e-10) and torch.allclose(d2, a2, atol=1e-10)
    planes = [f0 * s]
    if order >= 1:
        planes += [d1 * z[1 + k] * s for k in range(dims)]
    if order == 2:
        planes += [(d2 * z[1 + k] ** 2 + d1 * z[1 + dims + k]) * s for k in range(dims)]
    return torch.stack(planes)


CASES = [  # cj, dims, order, rows, rows_per_geom, k, n, act, use_escale, use_cvec
    (1, 0, 0, 300, 100, 10, 64, None, False, False),
    (1, 0, 0, 257, 0, 69, 96, 'silu', False, False),
    (4, 3, 1, 200, 50, 64, 384, 'silu', False, True),
    (3, 2, 1, 130, 65, 176, 352, 'silu', True, False),
    (7, 3, 2, 150, 75, 64, 128, 'silu', True, True),
    (5, 2, 2, 99, 33, 40, 3, 'tanh', True, False),
    (7, 3, 2, 64, 64, 3, 64, None, False, False),
    (4, 3, 1, 1000, 500, 384, 128, 'tanh', False, False),
    (7, 3, 2, 1536, 768, 64, 384, 'silu', True, True),
    (5, 2, 2, 2000, 1000, 176, 352, 'silu', True, False),
    (1, 0, 0, 4100, 0, 131, 128, 'silu', False, False),
    (4, 3, 1, 3000, 1500, 128, 4, 'silu', False, False),
    (3, 2, 1, 2048, 1024, 10, 64, None, False, False),
    # shapes served by the thin / small-rows kernels (jet_linear_thin.cu)
    (1, 0, 0, 32, 0, 1024, 384, None, False, False),      # per-geometry constant of the concat layer: rows = B
    (1, 0, 0, 20, 0, 100, 50, None, False, False),
    (1, 0, 0, 5000, 2500, 128, 4, 'silu', False, False),  # last layer, value only
    (4, 3, 1, 4096, 1024, 352, 3, 'silu', True, False),   # last layer behind a neural operator (branch scaling)
    (4, 3, 1, 4096, 2048, 3, 64, None, False, False),     # first layer, k = D
    (1, 0, 0, 6000, 0, 7, 64, None, False, False),        # first layer of a value-only encoder
    (5, 2, 2, 3000, 1500, 128, 8, 'tanh', False, True),
    # value-only layers on the tensor-memory-operand kernels (ws_fwd1.cu, ws_dw1.cuh): two output tiles with a ragged
    # second one, branch scaling + per-geometry constant, a contraction that ends inside a ring stage
    (1, 0, 0, 2048, 512, 96, 160, 'silu', True, True),
    (1, 0, 0, 3000, 0, 384, 128, 'tanh', False, False),
    (1, 0, 0, 1500, 0, 40, 300, 'silu', False, False),
]


@pytest.mark.parametrize('case', CASES)
def test_jet_linear_forward_backward(ops, case):
    cj, dims, order, rows, rpg, k, n, act, use_e, use_c = case
    g = torch.Generator().manual_seed(rows + k)
    n_geom = rows // rpg if rpg else 1
    zin = torch.randn(cj, rows, k, generator=g)
    w_full = torch.randn(n, k + 5, generator=g) / math.sqrt(k)     # the layer uses columns [2, 2+k) of a wider weight
    bias = torch.randn(n, generator=g)
    escale = torch.randn(n_geom, k, generator=g) if use_e else None
    cvec = torch.randn(n_geom, n, generator=g) if use_c else None
    gout = torch.randn(cj, rows, n, generator=g)

    # fp64 reference with autograd
    zr = zin.double().requires_grad_(True)
    wr = w_full.double().requires_grad_(True)
    br = bias.double().requires_grad_(True)
    er = escale.double().requires_grad_(True) if use_e else None
    cr = cvec.double().requires_grad_(True) if use_c else None
    geom = (torch.arange(rows) // rpg) if rpg else torch.zeros(rows, dtype=torch.long)
    s = er[geom] if use_e else torch.ones(rows, k, dtype=torch.float64)
    a = act_jet_ref(act, zr, dims, order, s)
    zo = a @ wr[:, 2:2 + k].t()
    add0 = cr[geom] if use_c else br[None, :]      # with cvec the caller folds the bias into it
    zo = torch.cat([(zo[0] + add0)[None], zo[1:]])
    (zo * gout.double()).sum().backward()

    from porous_cfd_b200.ops import Jet
    zj = Jet.empty(cj, rows, k, 'cuda'); zj.t[:, :, :k].copy_(zin)
    wd = dev(w_full)
    ed = dev(escale) if use_e else None      # must outlive tin (tin holds a raw device pointer)
    tin = ops.make_intrans(act, 0, ed) if (act or use_e) else None
    cv = dev(cvec) if use_c else None
    out = ops.jet_linear_fwd(zj, tin, wd, 2, k, None if use_c else dev(bias), cv, rpg, n)
    assert rel_l2(out.t[:, :, :n].cpu().double(), zo.detach()) < 5e-6

    gj = Jet.empty(cj, rows, n, 'cuda'); gj.t[:, :, :n].copy_(gout)
    ge = torch.zeros(n_geom, k, device='cuda') if use_e else None
    gzin = ops.jet_linear_bwd_dx(gj, wd, 2, zj, tin, ge, rpg, k, n)
    assert rel_l2(gzin.t[:, :, :k].cpu().double(), zr.grad) < GEMM_TOL
    if use_e:
        assert rel_l2(ge.cpu().double(), er.grad) < GEMM_TOL

    gw = torch.zeros_like(wd)
    gb = torch.zeros(n, device='cuda')
    gc = torch.zeros(n_geom, n, device='cuda') if use_c else None
    ws = torch.empty(ops.dw_workspace_bytes(cj, rows, rpg, k, n), dtype=torch.uint8, device='cuda')
    ops.jet_linear_bwd_dw(gj, zj, tin, gw, 2, None if use_c else gb, gc, rpg, k, n, ws)
    assert rel_l2(gw.cpu().double(), wr.grad) < GEMM_TOL
    assert float(gw[:, :2].abs().max()) == 0.0 and float(gw[:, 2 + k:].abs().max()) == 0.0   # neighbours untouched
    if use_c:
        assert rel_l2(gc.cpu().double(), cr.grad) < GEMM_TOL
    else:
        assert rel_l2(gb.cpu().double(), br.grad) < GEMM_TOL
    # accumulation semantics: a second call doubles the result
    ops.jet_linear_bwd_dw(gj, zj, tin, gw, 2, None if use_c else gb, gc, rpg, k, n, ws)
    assert rel_l2(gw.cpu().double(), 2 * wr.grad) < GEMM_TOL


def test_jet_linear_with_fallback_kernels():
    """The same layer cases with the tensor-memory-operand kernels and programmatic dependent launch switched off
    (PCFD_FWD1=0 PCFD_DW1=0 PCFD_PDL=0): the general engine-2 kernels stay correct for value-only layers."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PCFD_FWD1='0', PCFD_DW1='0', PCFD_PDL='0')
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-q', '-x', '-m', 'gpu', '-k',
                        'test_jet_linear_forward_backward or test_dropout_mask'], env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])


@pytest.mark.parametrize('cj', [4, 1])
def test_dropout_mask_is_consistent_between_passes(ops, cj):
    """The same (seed, salt) must give the same mask in forward, dX and dW; keep-rate ~ 1-p."""
    from porous_cfd_b200.ops import Jet
    rows, k, n, p = 512, 64, 32, 0.25
    g = torch.Generator().manual_seed(0)
    zj = Jet.empty(cj, rows, k, 'cuda'); zj.t.copy_(torch.randn(cj, rows, k, generator=g))
    w = torch.eye(k, device='cuda')[:n].contiguous()                 # zout = first n transformed inputs
    seed = torch.tensor([1234], dtype=torch.int64, device='cuda')
    tin = ops.make_intrans(None, 0, None, p, seed, salt=3)
    out = ops.jet_linear_fwd(zj, tin, w, 0, k, None, None, 0, n).t[:, :, :n]
    ratio = out / zj.t[:, :, :n]
    kept = ratio[0] != 0
    assert abs(float(kept.float().mean()) - (1 - p)) < 0.03
    assert torch.allclose(ratio[0][kept], torch.full_like(ratio[0][kept], 1 / (1 - p)), rtol=1e-6)
    for c in range(1, cj):
        assert torch.equal(ratio[c] != 0, kept)                      # one mask for all jet channels
    gj = Jet(torch.ones(cj, rows, n, device='cuda'), n)
    gz = ops.jet_linear_bwd_dx(gj, w, 0, zj, tin, None, 0, k, n).t[:, :, :n]
    assert torch.equal(gz[0] != 0, kept)
    # dW with gzout = 1: every row of gw is the column sum of the transformed input, i.e. of `out` where w is the identity
    gw = torch.zeros(n, k, device='cuda')
    ws = torch.empty(ops.dw_workspace_bytes(cj, rows, 0, k, n), dtype=torch.uint8, device='cuda')
    ops.jet_linear_bwd_dw(gj, zj, tin, gw, 0, None, None, 0, k, n, ws)
    want = out.double().sum(dim=(0, 1))
    assert rel_l2(gw[0, :n].cpu().double(), want.cpu()) < 1e-5
    # another step seed gives another mask
    ops.advance_seed(seed)
    out2 = ops.jet_linear_fwd(zj, tin, w, 0, k, None, None, 0, n).t[0, :, :n]
    assert not torch.equal(out2 != 0, kept)


# ---------------------------------------------------------------------------------------------
# ReLoBRaLo update kernel against the vectors of the reference's RelobraloScaler
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', ['rho1_b1', 'rho0_b3', 'rho1_b2_tau'])
def test_relobralo_kernel_matches_reference_vectors(ops, case):
    z = np.load(f'{GOLDEN}/relobralo.npz')
    n, alpha, beta, tau, bs, steps = z[f'{case}/meta']
    n, bs, steps = int(n), int(bs), int(steps)
    init, prev, lam = (torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda'), torch.ones(n, device='cuda'))
    step = torch.zeros(1, dtype=torch.int64, device='cuda')
    w = torch.ones(16, device='cuda')
    for s in range(steps):
        losses = dev(torch.from_numpy(z[f'{case}/losses'][s]))
        ops.relobralo_update(losses, n, init, prev, lam, step, bs, float(alpha), float(beta), float(tau), 1e-8, 8421, w)
        got = (w[:n] * losses).cpu()
        assert max_rel(got, torch.from_numpy(z[f'{case}/weighted'][s])) < 1e-5, (case, s)
    assert int(step.item()) == steps


def test_fused_adam_matches_torch(ops):
    """pcfd_adam_step against torch.optim.Adam (models/pipn/pipn_foam.py:102-105) on a flat buffer, 4 steps,
    learning rate changed in between (ExponentialLR)."""
    g = torch.Generator().manual_seed(5)
    n = 100003
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone().cuda())
    opt = torch.optim.Adam([ref], lr=1e-3)
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device='cuda')
    lr = torch.full((1,), 1e-3, device='cuda')
    for it in range(4):
        grad = torch.randn(n, generator=g).cuda() * (10.0 ** (it - 2))
        ref.grad = grad.clone()
        opt.step()
        ops.adam_step(p, grad * 4.0, m, v, step, lr, 0.9, 0.999, 1e-8, 0.25)     # grad_scale undoes the factor 4
        assert rel_l2(p.double().cpu(), ref.detach().double().cpu()) < 1e-6
        opt.param_groups[0]['lr'] *= 0.999
        lr.mul_(0.999)
    assert int(step.item()) == 4


# ---------------------------------------------------------------------------------------------
# sparse backward of (last MLP layer -> max pool) and the compacted backward of long segments
# ---------------------------------------------------------------------------------------------

POOL_CASES = [  # n_seg, seg_len, k, c, act_in, act_pool, masked slots
    (500, 17, 64, 128, 'silu', 'silu', True),       # set abstraction level 0 of config 2 (K + 1 = 17 slots)
    (130, 17, 128, 256, 'silu', 'silu', True),      # level 1
    (8, 125, 256, 1024, 'silu', 'silu', False),     # global set abstraction: medium segments, many channels
    (40, 65, 131, 128, None, 'silu', True),         # single-layer level (windbreaks SA1): raw edge features, odd k
    (64, 9, 6, 64, None, 'tanh', False),            # manufactured SA0 [6, 64]
    (3, 300, 40, 50, 'tanh', 'tanh', False),        # more rows than channels
]


@pytest.mark.parametrize('case', POOL_CASES)
def test_pool_layer_bwd_matches_autograd(ops, case):
    n_seg, L, k, c, act_in, act_pool, masked = case
    from porous_cfd_b200.ops import Jet
    g = torch.Generator().manual_seed(n_seg + L + k)
    rows = n_seg * L
    zin = torch.randn(rows, k, generator=g)
    w = torch.randn(c, k, generator=g) / math.sqrt(k)
    b = torch.randn(c, generator=g)
    gout = torch.randn(n_seg, c, generator=g)
    slots = None
    if masked:
        slots = torch.where(torch.rand(n_seg, L, generator=g) < 0.3, -1, 1).int()
        slots[:, 0] = 1                                   # a segment always keeps its self loop
    fa = {None: lambda x: x, 'silu': torch.nn.functional.silu, 'tanh': torch.tanh}
    zr = zin.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    br = b.double().requires_grad_(True)
    zo = fa[act_in](zr) @ wr.t() + br
    vals = fa[act_pool](zo).reshape(n_seg, L, c)
    if masked:
        vals = vals.masked_fill((slots < 0)[:, :, None], -float('inf'))
    ref_arg = vals.max(dim=1).indices

    zj = Jet.empty(1, rows, k, 'cuda'); zj.t[0, :, :k].copy_(zin)
    zj.t[0, :, k:] = float('nan')                          # row padding must never be read as data
    tin = ops.make_intrans(act_in) if act_in else None
    wd, bd = dev(w), dev(b)
    zl = ops.jet_linear_fwd(zj, tin, wd, 0, k, bd, None, 0, c)
    sd = dev(slots) if masked else None
    pooled, arg, zsel = ops.segmax_fwd_z(zl.t[0], act_pool, sd, n_seg, L, c)
    # the rows the kernel selected (two maxima within one fp32 ulp may swap against the fp64 restatement: rare, and then
    # either row is a correct arg-max of the fp32 forward; the gradient is checked for the rows actually selected)
    sel = arg.cpu().long()
    assert int((sel != ref_arg).sum()) <= 2
    out = vals.gather(1, sel[:, None, :]).squeeze(1)
    (out * gout.double()).sum().backward()
    assert rel_l2(pooled[:, :c].cpu().double(), out.detach()) < 5e-6
    assert ops.pool_layer_bwd_supported(n_seg, L, k, c, tin, zj.ld)
    gd = dev(gout)
    gw = torch.zeros_like(wd)
    gb = torch.zeros(c, device='cuda')
    ws = torch.empty(ops.pool_layer_bwd_workspace_bytes(n_seg, L, k, c), dtype=torch.uint8, device='cuda')
    gzin = ops.pool_layer_bwd(gd, gd.stride(0), arg, zsel, act_pool, n_seg, L, c, zj, tin, k, wd, gw, gb, True, ws)
    assert rel_l2(gw.cpu().double(), wr.grad) < GEMM_TOL
    assert rel_l2(gb.cpu().double(), br.grad) < GEMM_TOL
    assert rel_l2(gzin.t[0, :, :k].cpu().double(), zr.grad) < GEMM_TOL
    # accumulation semantics of the parameter gradients, and bit-identical repeat (fixed summation order)
    gw2 = torch.zeros_like(wd)
    ops.pool_layer_bwd(gd, gd.stride(0), arg, zsel, act_pool, n_seg, L, c, zj, tin, k, wd, gw2, None, False, ws)
    assert torch.equal(gw2, gw)
    ops.pool_layer_bwd(gd, gd.stride(0), arg, zsel, act_pool, n_seg, L, c, zj, tin, k, wd, gw2, None, False, ws)
    assert rel_l2(gw2.cpu().double(), 2 * wr.grad) < GEMM_TOL

    # the same gradient through the dense form the sparse one replaces
    gz = ops.segmax_bwd(gd, gd.stride(0), arg, zl.t[0], act_pool, n_seg, L, c)
    dense = ops.jet_linear_bwd_dx(Jet(gz, c), wd, 0, zj, tin, None, 0, k, c)
    assert rel_l2(gzin.t[0, :, :k].cpu().double(), dense.t[0, :, :k].cpu().double()) < GEMM_TOL


def test_pool_compact_rows_carry_the_whole_gradient(ops):
    """Long segments: the backward of a 2-layer encoder on the <= C selected rows per segment equals the dense one."""
    from porous_cfd_b200 import engine
    from porous_cfd_b200.ops import Jet
    n_seg, L, k0, k1, c = 4, 700, 24, 96, 80
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n_seg * L, k0, generator=g)
    lin0 = torch.nn.Linear(k0, k1)
    lin1 = torch.nn.Linear(k1, c)
    gout = torch.randn(n_seg, c, generator=g)
    xr = x.double()
    m0, m1 = lin0.double(), lin1.double()
    out = torch.nn.functional.silu(m1(torch.nn.functional.silu(m0(xr)))).reshape(n_seg, L, c).max(dim=1).values
    (out * gout.double()).sum().backward()
    want = {id(p): p.grad.clone() for p in (m0.weight, m0.bias, m1.weight, m1.bias)}
    lin0, lin1 = lin0.float().cuda(), lin1.float().cuda()
    for p in (lin0.weight, lin0.bias, lin1.weight, lin1.bias):
        p.grad = None
    ctx = engine.StepContext(torch.device('cuda'))
    params = [lin0.weight, lin0.bias, lin1.weight, lin1.bias]
    for p in params:
        ctx.grads[id(p)] = torch.zeros_like(p)
    layers, pending = engine.mlp_chain([lin0, lin1], 'silu', True)
    zj = Jet.empty(1, n_seg * L, k0, 'cuda'); zj.t[0, :, :k0].copy_(x)
    zs = engine.chain_forward(ctx, layers, zj, L)
    pooled, arg, zsel = ops.segmax_fwd_z(zs[-1].t[0], pending[0], None, n_seg, L, c)
    gd = dev(gout)
    assert not ops.pool_layer_bwd_supported(n_seg, L, k1, c, ops.make_intrans('silu'), zs[-2].ld)   # -> compaction path
    engine.pool_backward(ctx, layers, zs, gd, gd.stride(0), arg, zsel, pending[0], n_seg, L, rows_per_geom=L)
    torch.cuda.synchronize()
    ref = [m0.weight, m0.bias, m1.weight, m1.bias]
    for p, r in zip(params, ref):
        assert rel_l2(ctx.grads[id(p)].cpu().double(), want[id(r)]) < GEMM_TOL

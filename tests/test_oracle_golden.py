"""The oracle is pinned: it must reproduce the outputs the UNMODIFIED reference produced
(tests/golden/*.npz, written by tests/golden/make_golden.py in the build container)."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, TINY, TINY_SHAPE, flat, load_fixture, max_rel, rel_l2
from oracle import pinn_oracle, pyg_restate, ref_shim
from porous_cfd_b200 import synthetic

TOL = 2e-5  # fp32 CPU vs fp32 CPU, same op order: differences come only from thread-count dependent reductions


@pytest.mark.parametrize('name', TINY)
@pytest.mark.parametrize('mode', ['reference', 'true'])
def test_oracle_matches_reference_fixture(name, mode):
    spec = synthetic.model_spec(name)
    data, domain, params, out = load_fixture(name)
    labels = synthetic.build_labels(spec['layout'])
    got = pinn_oracle.step_with_grads(spec, params, data, labels, domain, laplacian=mode)
    ref = out[mode]
    assert max_rel(got['losses'], ref['losses']) < TOL
    assert abs(float(got['loss']) - float(ref['loss'])) / abs(float(ref['loss'])) < TOL
    keys = list(params)
    assert rel_l2(flat(got['grads'], keys), flat(ref['grads'], keys)) < TOL
    assert max_rel(got['u_error'], ref['u_error']) < TOL and max_rel(got['p_error'], ref['p_error']) < TOL


@pytest.mark.parametrize('name', TINY)
def test_synthetic_batch_is_reproducible(name):
    """The fixture inputs are regenerated bit for bit from the seed (same generator everywhere)."""
    spec = synthetic.model_spec(name)
    data, domain, _, out = load_fixture(name)
    d2, _, dom2 = synthetic.make_batch(spec['layout'], seed=out['seed'], **TINY_SHAPE)
    assert torch.equal(d2, data)
    for k in domain:
        assert torch.equal(dom2[k], domain[k])


def test_index_fixtures():
    z = np.load(f'{GOLDEN}/index_ops.npz')
    for tag in 'abc':
        nb, n, d, ratio, r, k = z[f'{tag}/meta']
        pos = torch.from_numpy(z[f'{tag}/pos'])
        batch = torch.arange(int(nb)).repeat_interleave(int(n))
        idx = pyg_restate.fps(pos, batch, float(ratio))
        assert torch.equal(idx, torch.from_numpy(z[f'{tag}/fps']))
        row, col = pyg_restate.radius(pos, pos[idx], float(r), batch, batch[idx], int(k))
        assert torch.equal(row, torch.from_numpy(z[f'{tag}/row'])) and torch.equal(col, torch.from_numpy(z[f'{tag}/col']))
        # properties: sorted by (row, col), at most K per row, every hit strictly inside the ball
        assert int(torch.bincount(row).max()) <= int(k)
        d2 = ((pos[col] - pos[idx][row]) ** 2).sum(1)
        assert bool((d2 < float(r) ** 2 + 1e-6).all())
    pos = torch.from_numpy(z['ties/pos'])
    assert torch.equal(pyg_restate.fps(pos, None, 0.5), torch.from_numpy(z['ties/fps']))
    assert z['ties/fps'].tolist() == [0, 1, 3]  # duplicates: the lowest index wins


def test_fps_general_path_equals_lockstep_path():
    g = torch.Generator().manual_seed(1)
    pos = torch.rand(3 * 50, 3, generator=g)
    batch = torch.arange(3).repeat_interleave(50)
    lock = pyg_restate.fps(pos, batch, 0.3)
    # ragged batch vector forces the per-element path; results on the shared elements must agree
    pos2 = torch.cat([pos, torch.rand(7, 3, generator=g)])
    batch2 = torch.cat([batch, torch.full((7,), 3)])
    gen = pyg_restate.fps(pos2, batch2, 0.3)
    assert torch.equal(gen[:lock.numel()], lock)


def test_manufactured_known_answer():
    """Exact analytic fields give a zero NS-Darcy residual through the oracle's loss restatement
    (KAT from examples/manufactured_solutions/manufactured_dataset.py:46-67)."""
    z = np.load(f'{GOLDEN}/manufactured_kat.npz')
    spec = synthetic.model_spec('manufactured_pipn_pp')
    labels = synthetic.build_labels('manufactured')
    data = torch.from_numpy(z['data']).double()
    internal = pinn_oracle.rows(data, torch.from_numpy(z['domain/internal']))
    t = lambda k: torch.from_numpy(z[k])
    res = pinn_oracle.momentum_residual(spec, internal, labels, t('u'), t('jac'), t('lap'), t('dp'))
    assert float(res.abs().max()) < 1e-6
    assert float(pinn_oracle.continuity_residual(spec, t('jac')).abs().max()) < 1e-12


@pytest.mark.skipif(not ref_shim.available(), reason='/root/reference only exists in the build container')
def test_oracle_matches_live_reference():
    """Where the reference tree is present, re-run it live (not just the stored fixture)."""
    import importlib.util
    import os
    spec_ = importlib.util.spec_from_file_location('make_golden', os.path.join(GOLDEN, 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mg)
    ref_shim.install()
    name = 'tiny_pigano'
    spec = synthetic.model_spec(name)
    torch.manual_seed(5)
    model = mg.build_reference_model(spec).eval()
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    data, labels, domain = synthetic.make_batch(spec['layout'], seed=99, **TINY_SHAPE)
    loss, losses, _, _, grads = mg.reference_step(model, spec, data, labels, domain, 'reference')
    got = pinn_oracle.step_with_grads(spec, params, data, labels, domain, 'reference')
    assert max_rel(got['losses'], losses) < TOL


@pytest.mark.parametrize('case', ['rho1_b1', 'rho0_b3', 'rho1_b2_tau'])
def test_relobralo_restatement_matches_reference_vectors(case):
    """oracle.Relobralo against the weighted loss vectors the reference's RelobraloScaler returned
    (tests/golden/relobralo.npz, written by tests/golden/make_relobralo_golden.py)."""
    z = np.load(f'{GOLDEN}/relobralo.npz')
    n, alpha, beta, tau, bs, steps = z[f'{case}/meta']
    sc = pinn_oracle.Relobralo(int(n), alpha=float(alpha), rho=float(beta), tau=float(tau), batch_size=int(bs))
    for s in range(int(steps)):
        got = sc(torch.from_numpy(z[f'{case}/losses'][s]))
        assert max_rel(got, torch.from_numpy(z[f'{case}/weighted'][s])) < 1e-6


# ---------------------------------------------------------------------------------------------
# host-side features of FoamDataset (dataset/foam_dataset.py:360-395): oracle/ingest_oracle.py against the vectors the
# unmodified reference produced (tests/golden/make_ingest_golden.py)
# ---------------------------------------------------------------------------------------------
INGEST_CASES = ['abc3d', 'duct2d_minmax', 'std3d_ragged']


@pytest.mark.parametrize('case', INGEST_CASES)
def test_ingest_oracle_matches_reference_vectors(case):
    import os
    from oracle import ingest_oracle
    z = np.load(os.path.join(GOLDEN, 'ingest.npz'))
    scale = z[f'{case}/coord_scale'] if z[f'{case}/coord_scale'].size else None
    si, sb = ingest_oracle.add_sdf(z[f'{case}/pos_internal'], z[f'{case}/pos_boundary'], z[f'{case}/region'], scale)
    assert np.allclose(si, z[f'{case}/sdf_internal'], rtol=1e-12, atol=1e-15)
    assert np.allclose(sb, z[f'{case}/sdf_boundary'], rtol=1e-12, atol=1e-15)
    # properties the reference's construction implies: boundary points have distance 0 to themselves, the largest
    # value is 1, porous-region points are negative
    assert float(np.abs(sb).max()) == 0.0 and abs(float(np.abs(si).max()) - 1.0) < 1e-12
    assert np.all(si[z[f'{case}/region'] > 0.5] <= 0) and np.all(si[z[f'{case}/region'] < 0.5] >= 0)
    ni = len(si)
    want = z[f'{case}/one_hot']
    cls = z[f'{case}/boundary_class']
    got = np.zeros_like(want)
    got[ni + np.arange(len(cls)), cls] = 1.0
    assert np.array_equal(got, want) and int(z[f'{case}/n_classes'][0]) == want.shape[1]


def test_ingest_oracle_class_order_and_collate():
    from oracle import ingest_oracle
    cats, cls = ingest_oracle.boundary_classes(['walls', 'inlet', 'walls', 'interface'])
    assert cats == ['inlet', 'interface', 'walls'] and cls.tolist() == [2, 0, 2, 1]
    oh = ingest_oracle.boundary_one_hot(['b', 'a'], 2)
    assert oh.tolist() == [[0, 0], [0, 0], [0, 1], [1, 0]]
    datas = [np.full((3, 2), i, dtype=np.float32) for i in range(4)]
    doms = [{'internal': np.arange(2) + i} for i in range(4)]
    d, dom = ingest_oracle.collate(datas, doms)
    assert d.shape == (4, 3, 2) and dom['internal'].tolist() == [[0, 1], [1, 2], [2, 3], [3, 4]]

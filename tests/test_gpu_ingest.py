"""Batch ingestion on the device (csrc/ingest.cu, dataset/device_dataset.py) against the vectors of the unmodified
reference (tests/golden/ingest.npz), the CPU oracle and the host collate_fn."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import ingest_oracle
from porous_cfd_b200 import synthetic
from porous_cfd_b200.dataset.device_dataset import DeviceFoamDataset
from porous_cfd_b200.dataset.foam_data import FoamData
from porous_cfd_b200.dataset.foam_dataset import collate_fn

pytestmark = pytest.mark.gpu

SDF_RTOL, SDF_ATOL = 2e-5, 2e-7      # float32 distances on the device against the reference's float64 (1e-4 budget)
CASES = ['abc3d', 'duct2d_minmax', 'std3d_ragged']


def frame(pos_i, pos_b, region, n_classes):
    """One geometry in the load_case row order: [C..., cellToRegion, sdf, boundaryId...]."""
    ni, nb, d = len(pos_i), len(pos_b), pos_i.shape[1]
    names = [f'C{a}' for a in 'xyz'[:d]] + ['cellToRegion', 'sdf'] + [f'boundaryId{i}' for i in range(n_classes)]
    labels = {n: None for n in names}
    labels['C'] = names[:d]
    labels['boundaryId'] = names[d + 2:]
    data = torch.full((ni + nb, len(names)), 7.0)                   # stale values must be overwritten
    data[:, :d] = torch.from_numpy(np.concatenate([pos_i, pos_b])).float()
    data[:ni, d] = torch.from_numpy(region).float()
    data[ni:, d] = 0.0
    domain = {'internal': torch.arange(ni), 'boundary': ni + torch.arange(nb)}
    return FoamData(data, labels, domain)


@pytest.mark.parametrize('case', CASES)
def test_features_match_reference_vectors(case):
    z = np.load(os.path.join(GOLDEN, 'ingest.npz'))
    nc = int(z[f'{case}/n_classes'][0])
    fd = frame(z[f'{case}/pos_internal'], z[f'{case}/pos_boundary'], z[f'{case}/region'], nc)
    ds = DeviceFoamDataset.from_samples([fd, fd])                    # two copies: the geometries must not interact
    scale = z[f'{case}/coord_scale']
    ds.add_sdf(torch.from_numpy(scale).float() if scale.size else None)
    ds.add_boundary_id(torch.from_numpy(z[f'{case}/boundary_class'])[None].repeat(2, 1))
    d = fd.data.shape[1] - 2 - nc
    want_sdf = np.concatenate([z[f'{case}/sdf_internal'], z[f'{case}/sdf_boundary']])
    for g in range(2):
        got = ds.data[g].cpu().double().numpy()
        assert np.allclose(got[:, d + 1], want_sdf, rtol=SDF_RTOL, atol=SDF_ATOL)
        assert np.array_equal(got[:, d + 2:], z[f'{case}/one_hot'])
        assert np.array_equal(got[:, :d + 1], fd.data[:, :d + 1].double().numpy())     # inputs untouched


def test_sdf_many_geometries_against_oracle():
    """Geometries of a batch are independent; more boundary points than one shared-memory tile (1024)."""
    g, ni, nb, d = 5, 1100, 1500, 3
    gen = torch.Generator().manual_seed(5)
    pos = torch.rand(g, ni + nb, d, generator=gen) * torch.tensor([1.0, 10.0, 0.1])
    region = (torch.rand(g, ni, generator=gen) < 0.4).float()
    samples = [frame(pos[i, :ni].numpy(), pos[i, ni:].numpy(), region[i].numpy(), 1) for i in range(g)]
    ds = DeviceFoamDataset.from_samples(samples)
    ds.add_sdf()
    for i in range(g):
        si, sb = ingest_oracle.add_sdf(pos[i, :ni].double().numpy(), pos[i, ni:].double().numpy(), region[i].numpy())
        got = ds.data[i, :, d + 1].cpu().double().numpy()
        assert np.allclose(got, np.concatenate([si, sb]), rtol=SDF_RTOL, atol=SDF_ATOL)
        assert abs(float(np.abs(got).max()) - 1.0) < 1e-6 and float(np.abs(got[ni:]).max()) == 0.0


def test_gather_blocks_single_tensor():
    from porous_cfd_b200 import ops
    gen = torch.Generator().manual_seed(2)
    ids = torch.tensor([5, 0, 5, 2], device='cuda')
    for shape, dtype in (((6, 37, 3), torch.float32), ((6, 16, 8), torch.float32), ((6, 9), torch.int64)):
        src = (torch.rand(shape, generator=gen) * 100).to(dtype).cuda()
        assert torch.equal(ops.gather_blocks(src, ids), src[ids])


@pytest.mark.parametrize('layout,n_obs', [('abc', 33), ('duct_variable', 33), ('manufactured', 0)])
def test_device_collate_is_bit_exact(layout, n_obs):
    data, labels, domain = synthetic.make_batch(layout, n_geometries=7, n_internal=150, n_boundary=101, n_obs=n_obs, seed=11)
    samples = [FoamData(data[i], labels, {k: v[i] for k, v in domain.items()}) for i in range(7)]
    ds = DeviceFoamDataset.from_samples(samples)
    assert len(ds) == 7
    for ids in ([3], [6, 0, 3, 3, 1], list(range(7))):
        want = collate_fn([samples[i] for i in ids])
        got = ds.batch(ids)
        assert got.labels is labels and torch.equal(got.data.cpu(), want.data)
        assert set(got.domain) == set(want.domain)
        for k in want.domain:
            assert got.domain[k].dtype == torch.int64 and torch.equal(got.domain[k].cpu(), want.domain[k])
    # an epoch of device batches covers every geometry once
    seen = torch.cat([b.data[:, 0, 0].cpu() for b in ds.batches(3, shuffle=True, generator=torch.Generator().manual_seed(0))])
    assert sorted(seen.tolist()) == sorted(data[:, 0, 0].tolist())
    with pytest.raises(IndexError):
        ds.batch([7])
    dev_ids = torch.tensor([2, 5], device='cuda')                    # ids already on the device: no host round trip
    assert torch.equal(ds.batch(dev_ids).data.cpu(), data[[2, 5]])


def test_training_step_from_a_device_batch():
    """The FoamData a DeviceFoamDataset hands out drives the training step exactly like the host-collated batch."""
    from porous_cfd_b200 import factory
    spec = synthetic.model_spec('tiny_pipn_pp')
    data, labels, domain = synthetic.make_batch(spec['layout'], n_geometries=5, n_internal=40, n_boundary=24, n_obs=10, seed=3)
    samples = [FoamData(data[i], labels, {k: v[i] for k, v in domain.items()}) for i in range(5)]
    torch.manual_seed(1)
    model = factory.build_model(spec).to('cuda').eval()
    ids = [4, 1]
    host = collate_fn([samples[i] for i in ids]).to('cuda')
    dev = DeviceFoamDataset.from_samples(samples).batch(ids)
    l_host = float(model.training_step(host, 0))
    l_dev = float(model.training_step(dev, 0))
    assert l_host == l_dev


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['tiny_pipn_pp', 'tiny_pigano_pp'])
def test_geometry_cache_reproduces_the_step(name):
    """SURVEY 8f rank 4, second half: FPS / ball-query results cached per geometry (the dataset is sampled once).  A
    batch drawn from the cache -- any geometries, any order -- brings exactly the centroids, edge slots and centroid
    positions that FPS + ball query + the self-loop rule give for that batch, and the training step (eager and from a
    CUDA graph) gives the same loss and gradients with it."""
    from helpers import load_fixture, rel_l2
    from porous_cfd_b200 import factory
    spec = synthetic.model_spec(name)
    _, _, params, _ = load_fixture(name)
    model = factory.build_model(spec)
    model.load_state_dict(params)
    model = model.cuda().eval()
    data, labels, domain = synthetic.make_batch(spec['layout'], 9, 40, 24, 10, seed=4)
    plain = DeviceFoamDataset(data, labels, domain)
    cached = DeviceFoamDataset(data, labels, domain)
    cached.build_geometry_cache(model, chunk=4)          # chunks of 4, 4, 1
    ex = model.executor
    for ids in ([3, 0, 7], [8, 8, 1, 2], [5]):
        want_batch = plain.batch(ids)
        got_batch = cached.batch(ids)
        assert want_batch.geometry is None and got_batch.geometry is not None
        want = ex.geometry(want_batch.data, labels, want_batch.domain)
        for lw, lg in zip(want, got_batch.geometry):
            for k in ('idx', 'slots', 'newpos'):
                assert torch.equal(lw[k], lg[k]), (ids, k)
        r0 = model.fused_step(want_batch)
        g0 = ex.flat_grad.clone()
        r1 = model.fused_step(got_batch)
        assert torch.equal(r0.out, r1.out) or float((r0.out - r1.out).abs().max()) <= 1e-6 * float(r0.out.abs().max())
        assert rel_l2(ex.flat_grad.double().cpu(), g0.double().cpu()) < 1e-6
    # the graphed public path: eager call, capture, replays, all with cached geometry
    model.cuda_graph = True
    for step, ids in enumerate(([1, 2, 3], [4, 5, 6], [7, 8, 0], [2, 2, 5])):
        batch = cached.batch(ids)
        loss = model.training_step(batch, step)
        ref = model.executor.step(batch.data, labels, batch.domain, model.laplacian)
        assert abs(float(loss) - float(ref.loss)) <= 1e-6 * abs(float(ref.loss)), step

"""common.training on the GPU: one epoch of train() on a synthetic dataset, checkpoint / resume of the full trainer
state, gradient accumulation, and -- on a box with >= 2 GPUs -- N-rank NCCL gradients == 1-rank gradients on the
concatenated batch (SURVEY.md section 4 item v; reference loop common/training.py:50-85)."""
import os
import socket
from argparse import Namespace

import pytest
import torch

import pcfd_import

pcfd_import.load()
from helpers import rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu


class _SyntheticSet(torch.utils.data.Dataset):
    """Per-geometry FoamData items, the item type of the reference's FoamDataset (dataset/foam_dataset.py:440-447)."""

    def __init__(self, layout, n, ni, nb, no, seed=5):
        from porous_cfd_b200 import synthetic
        from porous_cfd_b200.dataset.foam_data import FoamData
        data, labels, domain = synthetic.make_batch(layout, n, ni, nb, no, seed=seed)
        self.items = [FoamData(data[i], labels, {k: v[i] for k, v in domain.items()}) for i in range(n)]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def _args(tmp_path, **kw):
    base = dict(n_internal=40, n_boundary=24, n_observations=10, batch_size=4, precision='32', epochs=2, logs_dir=str(tmp_path),
                train_dir='', val_dir='', model='tiny', name='run', checkpoint=None, loss_scaler='fixed')
    base.update(kw)
    return Namespace(**base)


def _model(name='tiny_pigano', seed=3):
    from porous_cfd_b200 import factory, synthetic
    spec = synthetic.model_spec(name)
    torch.manual_seed(seed)
    return factory.build_model(spec), spec


def test_train_runs_epochs_and_writes_the_reference_artefacts(tmp_path):
    from porous_cfd_b200.common.training import train
    model, spec = _model()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    ds = _SyntheticSet(spec['layout'], 10, 40, 24, 10)      # 10 samples, batch 4: last batch is smaller
    hist = train(_args(tmp_path), model, ds, _SyntheticSet(spec['layout'], 3, 40, 24, 10, seed=9))
    assert len(hist) == 2 and all(torch.isfinite(torch.tensor(hist)))
    run = tmp_path / 'lightning_logs' / 'run'
    assert (run / 'model_meta.json').exists() and (run / 'model.ckpt').exists()
    ck = torch.load(run / 'model.ckpt', map_location='cpu')
    assert set(ck) >= {'state_dict', 'trainer', 'epoch', 'global_step'} and ck['global_step'] == 6 and ck['epoch'] == 2
    assert int(ck['trainer']['step']) == 6
    changed = [k for k, v in model.state_dict().items() if not torch.equal(v.cpu(), before[k])]
    assert len(changed) == len(before)


def test_resume_continues_exactly_where_the_run_stopped(tmp_path):
    """2 epochs in one go == 1 epoch, checkpoint, resume for the second (weights, Adam moments, step, lr, epoch, dropout
    counter)."""
    from porous_cfd_b200.common.training import train
    _, spec = _model()
    ds = _SyntheticSet(spec['layout'], 8, 40, 24, 10)
    val = _SyntheticSet(spec['layout'], 2, 40, 24, 10, seed=9)
    full, _ = _model()
    train(_args(tmp_path / 'a', epochs=2), full, ds, val)
    first, _ = _model()
    train(_args(tmp_path / 'b', epochs=1), first, ds, val)
    second, _ = _model(seed=11)                       # different init: everything must come from the checkpoint
    train(_args(tmp_path / 'c', epochs=2, checkpoint=str(tmp_path / 'b' / 'lightning_logs' / 'run' / 'model.ckpt')), second, ds, val)
    a = torch.load(tmp_path / 'a' / 'lightning_logs' / 'run' / 'model.ckpt', map_location='cpu')
    c = torch.load(tmp_path / 'c' / 'lightning_logs' / 'run' / 'model.ckpt', map_location='cpu')
    for k in a['state_dict']:
        # not bit for bit: the branch-scaling gradient is summed with atomics (order varies run to run)
        assert rel_l2(c['state_dict'][k].double(), a['state_dict'][k].double()) < 1e-4, k
    assert int(c['trainer']['step']) == int(a['trainer']['step'])
    assert rel_l2(c['trainer']['exp_avg'].double(), a['trainer']['exp_avg'].double()) < 1e-4


def test_train_with_relobralo_resumes_the_scaler_state(tmp_path):
    """train() with the adaptive scaler: the batch size reaches the scaler (the reference reads it from the trainer's
    DataLoader), the step counter follows the global step, and a resumed run keeps the restored init / previous losses
    (reference: RelobraloScaler keys on model.global_step, models/losses.py:93-124)."""
    from porous_cfd_b200.common.training import train
    from porous_cfd_b200.models.losses import RelobraloScaler
    model, spec = _model('tiny_pipn_pp')
    n_terms = 2 * spec['dims'] + 2 + spec['dims'] + 1
    model.loss_scaler = RelobraloScaler(n_terms, alpha=0.9, beta=1.0)
    ds = _SyntheticSet(spec['layout'], 8, 40, 24, 10)
    val = _SyntheticSet(spec['layout'], 2, 40, 24, 10, seed=9)
    hist = train(_args(tmp_path / 'a', epochs=2, loss_scaler='relobralo'), model, ds, val)
    assert all(torch.isfinite(torch.tensor(hist)))
    assert model.loss_scaler.batch_size == 4
    step, _ = model.loss_scaler.device_state(torch.device('cuda', 0))
    assert int(step) == 4                                   # 2 epochs x 2 steps
    init = model.loss_scaler.init_losses.clone()
    assert float(init.abs().sum()) > 0
    ck = tmp_path / 'a' / 'lightning_logs' / 'run' / 'model.ckpt'
    again, _ = _model('tiny_pipn_pp', seed=11)
    again.loss_scaler = RelobraloScaler(n_terms, alpha=0.9, beta=1.0)
    train(_args(tmp_path / 'b', epochs=3, loss_scaler='relobralo', checkpoint=str(ck)), again, ds, val)
    assert torch.equal(again.loss_scaler.init_losses.cpu(), init.cpu())       # not re-initialised by the first resumed step
    step2, _ = again.loss_scaler.device_state(torch.device('cuda', 0))
    assert int(step2) == 6


@pytest.mark.parametrize('name', ['tiny_pigano', 'tiny_pipn_pp'])
def test_train_from_an_hbm_resident_dataset_equals_the_host_loader(tmp_path, name):
    """train(args.device_dataset=True [, geometry_cache=True]): same sample order, same steps, same weights as the host
    DataLoader path (collation on the device replaces DataLoader + collate_fn + upload)."""
    from porous_cfd_b200.common.training import train
    _, spec = _model(name)
    ds = _SyntheticSet(spec['layout'], 10, 40, 24, 10)
    val = _SyntheticSet(spec['layout'], 2, 40, 24, 10, seed=9)
    host, _ = _model(name)
    train(_args(tmp_path / 'h', epochs=2), host, ds, val)
    dev, _ = _model(name)
    train(_args(tmp_path / 'd', epochs=2, device_dataset=True, geometry_cache=True), dev, ds, val)
    a = torch.load(tmp_path / 'h' / 'lightning_logs' / 'run' / 'model.ckpt', map_location='cpu')['state_dict']
    b = torch.load(tmp_path / 'd' / 'lightning_logs' / 'run' / 'model.ckpt', map_location='cpu')['state_dict']
    for k in a:
        assert rel_l2(b[k].double(), a[k].double()) < 1e-4, k


def test_accumulated_micro_batches_equal_one_batch():
    from porous_cfd_b200 import synthetic
    from porous_cfd_b200.dataset.foam_data import FoamData
    model, spec = _model()
    model = model.cuda().eval()
    data, labels, domain = synthetic.make_batch(spec['layout'], 4, 40, 24, 10, seed=2)
    whole = FoamData(data, labels, domain).to('cuda')
    model.fused_step(whole)
    g_whole = model.executor.flat_grad.clone()
    for i in range(2):
        part = FoamData(data[2 * i:2 * i + 2].contiguous(), labels, {k: v[2 * i:2 * i + 2].contiguous() for k, v in domain.items()}).to('cuda')
        model.fused_step(part, accumulate=i > 0)
    assert rel_l2((model.executor.flat_grad / 2).double().cpu(), g_whole.double().cpu()) < 1e-5


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out_dir, mode):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, 'tests'))
    import pcfd_import as pi
    pi.load()
    import torch.distributed as dist
    from porous_cfd_b200 import factory, synthetic
    from porous_cfd_b200.common.training import FlatAdamTrainer, shard_batch
    from porous_cfd_b200.dataset.foam_data import FoamData
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    os.environ['PCFD_DP_MULTIMEM'] = '0' if mode == 'fused_p2p' else '1'
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    spec = synthetic.model_spec('tiny_pigano')
    torch.manual_seed(3 + rank)                      # different replicas: the trainer's broadcast must align them
    model = factory.build_model(spec).cuda().eval()
    trainer = FlatAdamTrainer(model, fused_dp=mode != 'nccl')
    assert (trainer.dp is not None) == (mode != 'nccl'), 'symmetric memory was expected to be available'
    if trainer.dp is not None:
        assert trainer.dp['multimem'] == (mode == 'fused'), 'multicast was expected to be available'
    p0 = trainer.flat_param.clone().cpu()
    g = None
    for step in range(3):
        data, labels, domain = synthetic.make_batch(spec['layout'], 2 * world, 40, 24, 10, seed=21 + step)
        mine = shard_batch(FoamData(data, labels, domain), rank, world).to('cuda')
        res = model.fused_step(mine)
        if step == 0 and mode == 'nccl':
            trainer.reduce_gradients()
            g = (model.executor.flat_grad / world).cpu()
            trainer.step()
        else:
            trainer.train_step  # noqa: B018  (documented entry point; the two calls below are what it does)
            trainer.reduce_gradients()
            trainer.step()
    torch.cuda.synchronize()
    err = int(trainer.dp['epoch'][2]) if trainer.dp is not None else 0
    torch.save({'g': g, 'p0': p0, 'p1': trainer.flat_param.cpu(), 'loss': res.losses.cpu(), 'err': err},
               os.path.join(out_dir, f'{mode}_r{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('mode', ['nccl', 'fused', 'fused_p2p'])
def test_two_rank_step_equals_one_rank_on_the_concatenated_batch(tmp_path, mode):
    """Three optimizer steps on 2 ranks == the same steps in one process on the concatenated batches, for the NCCL tail
    (all-reduce + Adam) and for the one-kernel tail over peer memory (pcfd_dp_adam_step) with multimem (NVSwitch
    reduction / multicast) and with plain peer loads / stores."""
    import torch.multiprocessing as mp
    from porous_cfd_b200 import synthetic
    from porous_cfd_b200.common.training import FlatAdamTrainer
    from porous_cfd_b200.dataset.foam_data import FoamData
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    r = [torch.load(tmp_path / f'{mode}_r{i}.pt') for i in range(world)]
    assert r[0]['err'] == 0 and r[1]['err'] == 0, 'a cross-rank barrier timed out'
    assert torch.equal(r[0]['p0'], r[1]['p0']), 'replicas were not aligned by the broadcast'
    assert torch.equal(r[0]['p1'], r[1]['p1']), 'replicas diverged'
    # the same steps in one process on all 2*world geometries, from the broadcast weights
    model, spec = _model()
    model = model.cuda().eval()
    trainer = FlatAdamTrainer(model)
    trainer.flat_param.copy_(r[0]['p0'].cuda())
    for step in range(3):
        data, labels, domain = synthetic.make_batch(spec['layout'], 2 * world, 40, 24, 10, seed=21 + step)
        model.fused_step(FoamData(data, labels, domain).to('cuda'))
        if step == 0 and r[0]['g'] is not None:
            assert rel_l2(r[0]['g'].double(), model.executor.flat_grad.double().cpu()) < 1e-5
        trainer.step()
    assert rel_l2(r[0]['p1'].double(), trainer.flat_param.double().cpu()) < 1e-5

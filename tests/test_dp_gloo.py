"""Data-parallel shell on CPU (gloo, world size 2): sharding a batch over ranks, summing the flat
gradient with one all-reduce and dividing by the world size gives the single-process result on
the whole batch.  The per-rank "step" here is the CPU oracle (the CUDA step needs a GPU); what is
under test is the host logic of porous_cfd_b200.common.training (shard_batch + reduction rule)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import flat, rel_l2


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, 'tests'))
    import pcfd_import
    pcfd_import.load()
    from oracle import pinn_oracle
    from porous_cfd_b200 import factory, synthetic
    from porous_cfd_b200.common.training import shard_batch
    from porous_cfd_b200.dataset.foam_data import FoamData

    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    spec = synthetic.model_spec('tiny_pigano')
    torch.manual_seed(3)
    params = {k: v.detach().clone() for k, v in factory.build_model(spec).state_dict().items()}
    data, labels, domain = synthetic.make_batch(spec['layout'], 4, 24, 16, 6, seed=21)
    whole = FoamData(data, labels, domain)
    mine = shard_batch(whole, rank, world)
    assert mine.data.shape[0] == 2 and torch.equal(mine.data, data[2 * rank:2 * rank + 2])
    got = pinn_oracle.step_with_grads(spec, params, mine.data, labels, mine.domain, 'reference')
    keys = list(params)
    g = flat(got['grads'], keys).float()
    losses = got['losses'].detach().clone()
    # the reduction rule of FlatAdamTrainer.reduce_gradients: sum, then 1/world
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    g.mul_(1.0 / world)
    dist.all_reduce(losses, op=dist.ReduceOp.SUM)
    losses.mul_(1.0 / world)
    if rank == 0:
        ref = pinn_oracle.step_with_grads(spec, params, data, labels, domain, 'reference')
        torch.save({'g': g, 'ref': flat(ref['grads'], keys).float(), 'losses': losses,
                    'ref_losses': ref['losses'].detach()}, os.path.join(out_dir, 'dp.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = torch.load(os.path.join(tmp_path, 'dp.pt'))
    assert rel_l2(r['g'].double(), r['ref'].double()) < 1e-5
    assert float(((r['losses'] - r['ref_losses']).abs() / r['ref_losses'].abs()).max()) < 1e-5


def test_shard_batch_rejects_uneven_split():
    import pcfd_import
    pcfd_import.load()
    from porous_cfd_b200 import synthetic
    from porous_cfd_b200.common.training import shard_batch
    from porous_cfd_b200.dataset.foam_data import FoamData
    data, labels, domain = synthetic.make_batch('duct_variable', 3, 8, 8, 2)
    with pytest.raises(ValueError):
        shard_batch(FoamData(data, labels, domain), 0, 2)

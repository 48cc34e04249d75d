"""Host logic of porous_cfd_b200.common.training that needs no GPU: the per-rank sample order (DistributedSampler rule,
what Lightning puts behind the reference's DataLoader, common/training.py:57) and the ReLoBRaLo resume rule."""
import torch

import pcfd_import

pcfd_import.load()
from porous_cfd_b200.common.training import epoch_indices  # noqa: E402
from porous_cfd_b200.models.losses import RelobraloScaler  # noqa: E402


def test_epoch_indices_match_torch_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler
    data = list(range(13))
    for world in (1, 2, 4, 8):
        for epoch in (0, 3):
            for rank in range(world):
                ref = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=True, seed=8421)
                ref.set_epoch(epoch)
                assert epoch_indices(13, epoch, rank, world) == list(iter(ref))


def test_every_rank_draws_the_same_number_of_samples_and_all_are_covered():
    for n in (13, 64, 5):
        for world in (2, 4, 8):
            parts = [epoch_indices(n, 1, r, world) for r in range(world)]
            assert len({len(p) for p in parts}) == 1
            assert set(sum(parts, [])) == set(range(n))
    assert epoch_indices(7, 0, 0, 1, shuffle=False) == list(range(7))


def test_relobralo_restored_buffers_are_not_reinitialised():
    a = RelobraloScaler(5)
    a.init_losses.copy_(torch.arange(1.0, 6.0))
    a.prev_losses.copy_(torch.arange(2.0, 7.0))
    b = RelobraloScaler(5)
    assert b._resume_step == 0
    b.load_state_dict(a.state_dict(), strict=True)       # reference key names only: no extra state
    assert b._resume_step >= 1                            # the kernel's step-0 branch must not run again
    b.set_global_step(40)
    assert b._resume_step == 40
    fresh = RelobraloScaler(5)
    fresh.load_state_dict(RelobraloScaler(5).state_dict())
    assert fresh._resume_step == 0                        # an untrained state keeps the first-step initialisation


def test_models_are_freed_by_reference_counting():
    """No reference cycle through the loss loggers: a dropped model is released at once, not when the cyclic collector
    gets to it (on the GPU a cyclic model would keep its CUDA graphs and buffers alive until then)."""
    import gc
    import weakref
    from porous_cfd_b200 import factory, synthetic
    was = gc.isenabled()
    gc.disable()
    try:
        for name in ('tiny_pipn_pp', 'tiny_pigano', 'tiny_pipn', 'tiny_pigano_pp', 'tiny_manufactured_pp'):
            model = factory.build_model(synthetic.model_spec(name))
            ref = weakref.ref(model)
            del model
            assert ref() is None, name
    finally:
        if was:
            gc.enable()

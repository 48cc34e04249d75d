"""Training driver: the reference's `train(args, model, train_data, val_data)` and
`build_arg_parser()` (reference common/training.py:21-85) without Lightning.

The reference hands the loop to `lightning.Trainer`, which silently becomes DDP when several GPUs
are visible.  Here the data-parallel shell is explicit: one process per GPU
(torchrun), every rank runs the fused step on its shard of the geometries, the flat fp32 gradient
is summed with ONE NCCL all-reduce and divided by the world size (DDP's average), then a fused Adam
updates the flat parameter buffer.  Losses are means over the local shard, so equal shards give the
same update as one process on the concatenated batch.
"""
from __future__ import annotations

import argparse
import json
import os
from argparse import ArgumentParser, Namespace
from pathlib import Path
from typing import Optional

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Dataset

from ..dataset.foam_data import FoamData
from ..dataset.foam_dataset import collate_fn


def get_log_steps(n_data, batch_size):
    return (n_data // batch_size) + min(1, n_data % batch_size)


def build_arg_parser() -> ArgumentParser:
    """Same flags and defaults as the reference (common/training.py:21-47)."""
    p = argparse.ArgumentParser()
    p.add_argument('--n-internal', type=int, default=1000, help='number of internal points to sample')
    p.add_argument('--n-boundary', type=int, default=200, help='number of boundary points to sample')
    p.add_argument('--n-observations', type=int, default=500, help='number of observation points to sample')
    p.add_argument('--batch-size', type=int, default=13)
    p.add_argument('--precision', type=str, default='bf16-mixed',
                   help='accepted for compatibility; the CUDA path computes in fp32 (3xTF32 on tensor cores)')
    p.add_argument('--epochs', type=int, default=3000)
    p.add_argument('--logs-dir', type=str, default=os.getcwd())
    p.add_argument('--train-dir', type=str, default='data/train')
    p.add_argument('--val-dir', type=str, default='data/val')
    p.add_argument('--model', type=str, help='model type. The available models depend on the experiment')
    p.add_argument('--name', type=str, default=None)
    p.add_argument('--checkpoint', type=str, default=None)
    p.add_argument('--loss-scaler', type=str, default='fixed')
    return p


class FlatAdamTrainer:
    """Fused step + gradient all-reduce + fused Adam on flat buffers.

    Multi-GPU: ONE all-reduce of the flat fp32 gradient (3.4-6 MB), then the Adam kernel with the 1/world of DDP's
    average folded in.  Both are graph-capturable (`dist.all_reduce` on NCCL is captured as a kernel node), so a
    data-parallel step can replay as one graph: `reduce_gradients(); step()` inside the capture, as bench.py does.
    Bucketing the reduction per chain was measured not to pay on this path: the tensor-core kernels of a step own the
    whole GPU one after the other, the weight-gradient queue drains at the very end of the reverse pass, and the only
    gradients that are final early (the last two decoder layers) are 6 % of the buffer."""

    def __init__(self, model, process_group=None, fused_dp: Optional[bool] = None):
        """`fused_dp`: at world > 1 use the one-kernel optimizer tail over peer memory (pcfd_dp_adam_step) instead of
        NCCL all-reduce + Adam.  Default: on when torch's symmetric memory is available for the group (PCFD_FUSED_DP=0
        switches it off); False forces the NCCL path (the independent implementation the fused kernel is tested against)."""
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        ex = model.executor
        # re-point every parameter into one flat buffer (same order as the executor's flat gradient) so that Adam is a
        # single fused launch
        total = ex.flat_grad.numel()
        self.dp = None
        if fused_dp is None:
            fused_dp = os.environ.get('PCFD_FUSED_DP', '1') != '0'
        if self.world > 1 and fused_dp:
            try:
                self._setup_symmetric(ex, total)
            except Exception as exc:      # no symmetric memory on this box / build: NCCL all-reduce + Adam
                import warnings
                warnings.warn(f'fused data-parallel optimizer tail unavailable ({type(exc).__name__}: {exc}); using NCCL')
                self.dp = None
        if self.dp is None:
            self.flat_param = torch.empty(total, dtype=torch.float32, device=ex.device)
        off = 0
        for p in ex.params:
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p)
            off += n
        ex.reset_graphs()    # captured step graphs hold the parameters' old addresses
        if self.world > 1:   # identical replicas
            dist.broadcast(self.flat_param, 0, group=self.group)
        opt_cfg = model.configure_optimizers()[0][0]
        g = opt_cfg.param_groups[0]
        self.betas, self.eps = g['betas'], g['eps']
        self.gamma = model.configure_optimizers()[1][0]['scheduler'].gamma
        # Adam state on the device (step and learning rate included): the update is one kernel, graph-capturable
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=ex.device)
        self.lr_dev = torch.full((1,), float(g['lr']), dtype=torch.float32, device=ex.device)
        self.optimizer = self     # `.optimizer.step()` of the earlier torch.optim-based trainer keeps working
        self.accumulate = 1       # micro-batches per optimizer step (their gradients are summed; 1/accumulate in Adam)
        if self.dp is not None:   # moments padded like the symmetric buffers (float4 units)
            pad = self.dp['n_pad'] - total
            if pad:
                self.exp_avg = torch.zeros(self.dp['n_pad'], dtype=torch.float32, device=ex.device)[:total]
                self.exp_avg_sq = torch.zeros(self.dp['n_pad'], dtype=torch.float32, device=ex.device)[:total]

    def _setup_symmetric(self, ex, total: int) -> None:
        """Flat gradient, flat parameters and the barrier flags in symmetric memory; peers' addresses into a DpPeers."""
        import torch.distributed._symmetric_memory as symm_mem
        from .. import _lib
        group = self.group if self.group is not None else dist.group.WORLD
        name = group.group_name
        n_pad = (total + 3) // 4 * 4
        dev = ex.device
        g = symm_mem.empty(n_pad, dtype=torch.float32, device=dev)
        p = symm_mem.empty(n_pad, dtype=torch.float32, device=dev)
        f = symm_mem.empty(int(_lib.load().pcfd_dp_flags_len()), dtype=torch.int32, device=dev)
        g.zero_()
        p.zero_()
        f.zero_()
        hg, hp, hf = (symm_mem.rendezvous(t, name) for t in (g, p, f))
        peers = _lib.DpPeers()
        for r in range(self.world):
            peers.grad[r], peers.param[r], peers.flags[r] = hg.buffer_ptrs[r], hp.buffer_ptrs[r], hf.buffer_ptrs[r]
        use_mc = os.environ.get('PCFD_DP_MULTIMEM', '1') != '0' and hg.multicast_ptr and hp.multicast_ptr
        peers.grad_mc = hg.multicast_ptr if use_mc else None
        peers.param_mc = hp.multicast_ptr if use_mc else None
        peers.rank, peers.world = hg.rank, self.world
        self.flat_param = p[:total]
        ex.rebind_flat_grad(g[:total])
        torch.cuda.synchronize()
        hf.barrier()              # every rank's buffers are zeroed before any peer touches them
        self.dp = {'peers': peers, 'handles': (hg, hp, hf), 'buffers': (g, p, f), 'n_pad': n_pad, 'multimem': bool(use_mc),
                   'epoch': torch.zeros(4, dtype=torch.int32, device=dev)}

    def reduce_gradients(self, grad: Optional[torch.Tensor] = None):
        """All-reduce of the flat gradient.  With the fused tail the reduction happens inside step(): nothing to do here
        for the executor's own gradient buffer."""
        if self.world > 1 and not (self.dp is not None and grad is None):
            dist.all_reduce(self.model.executor.flat_grad if grad is None else grad, op=dist.ReduceOp.SUM, group=self.group)

    def step(self, grad: Optional[torch.Tensor] = None):
        """Adam (torch.optim.Adam semantics) on the flat buffers, one kernel; 1/world of the all-reduce is folded in.
        `grad`: a flat gradient other than the executor's own buffer (after `loss.backward()` through the autograd
        seam that is `model.executor.last_flat_grad`, of which every `p.grad` is a view)."""
        from .. import ops
        scale = 1.0 / (self.world * self.accumulate)
        if self.dp is not None and grad is None:
            # reduce-scatter + Adam on this rank's slice + all-gather of the parameters, one kernel over peer memory
            ops.dp_adam_step(self.dp['peers'], self.exp_avg, self.exp_avg_sq, self.step_dev, self.lr_dev, self.betas[0],
                             self.betas[1], self.eps, scale, self.flat_param.numel(), self.dp['epoch'])
            return
        g = self.model.executor.flat_grad if grad is None else grad
        ops.adam_step(self.flat_param, g, self.exp_avg, self.exp_avg_sq, self.step_dev,
                      self.lr_dev, self.betas[0], self.betas[1], self.eps, scale)

    def train_step(self, batch: FoamData, laplacian: Optional[str] = None, graphed: bool = False):
        """One optimizer step on `batch` (already on the device).  `graphed`: replay the fused step from the executor's
        CUDA graph for this batch signature (first call eager, second captures) -- what `train` uses, because launching
        the ~110 kernels of a step from Python takes longer than the GPU needs to run them."""
        if graphed:
            ex = self.model.executor
            res = ex.graphed_step(batch.data, batch.labels, batch.domain, laplacian or self.model.laplacian,
                                  geo=getattr(batch, 'geometry', None))
        else:
            res = self.model.fused_step(batch, laplacian)
        self.reduce_gradients()
        self.step()
        return res

    def end_epoch(self):
        """ExponentialLR, interval = epoch (reference models/pipn/pipn_foam.py:102-105)."""
        self.lr_dev.mul_(self.gamma)

    # ---- trainer state (what Lightning's checkpoint carries beside the weights) ---------------------
    def state_dict(self) -> dict:
        return {'exp_avg': self.exp_avg.clone(), 'exp_avg_sq': self.exp_avg_sq.clone(), 'step': self.step_dev.clone(),
                'lr': self.lr_dev.clone(), 'dropout_seed': self.model.executor.ctx.seed_dev.clone()}

    def load_state_dict(self, state: dict) -> None:
        self.exp_avg.copy_(state['exp_avg'])
        self.exp_avg_sq.copy_(state['exp_avg_sq'])
        self.step_dev.copy_(state['step'])
        self.lr_dev.copy_(state['lr'])
        if 'dropout_seed' in state:       # the counter behind the dropout masks: a resumed run draws the masks it would have
            self.model.executor.ctx.seed_dev.copy_(state['dropout_seed'])


def shard_batch(batch: FoamData, rank: int, world: int) -> FoamData:
    """This rank's equal share of the geometries of an already collated batch (bench / tests; `train` shards the
    dataset indices instead, see epoch_indices)."""
    b = batch.data.shape[0]
    if b % world != 0:
        raise ValueError(f'batch of {b} geometries does not split evenly over {world} ranks')
    per = b // world
    sl = slice(rank * per, (rank + 1) * per)
    return FoamData(batch.data[sl].contiguous(), batch.labels, {k: v[sl].contiguous() for k, v in batch.domain.items()})


def epoch_indices(n: int, epoch: int, rank: int, world: int, shuffle: bool = True, seed: int = 8421) -> list:
    """The sample indices rank `rank` visits in `epoch`: torch's DistributedSampler rule, which is what Lightning puts
    behind the reference's DataLoader under DDP -- one permutation per epoch shared by all ranks (seed + epoch), padded
    by wrap-around to a multiple of the world size, rank r takes every world-th index.  Every rank gets the same number
    of samples, so no batch ever fails to split."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g).tolist()
    else:
        idx = list(range(n))
    if world > 1:
        total = (n + world - 1) // world * world
        idx = (idx + idx[:total - n]) if total > n else idx
        idx = idx[rank:total:world]
    return idx


class _EpochSampler(torch.utils.data.Sampler):
    def __init__(self, n, rank, world, shuffle):
        self.n, self.rank, self.world, self.shuffle, self.epoch = n, rank, world, shuffle, 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __iter__(self):
        return iter(epoch_indices(self.n, self.epoch, self.rank, self.world, self.shuffle))

    def __len__(self):
        return (self.n + self.world - 1) // self.world


def train(args: Namespace, model, train_data: Dataset, val_data: Dataset):
    """Train `model`; writes model_meta.json and model.ckpt under logs_dir/lightning_logs/<name>
    like the reference.  Under torchrun every rank draws `batch_size` geometries per step from its shard of the epoch
    (DistributedSampler semantics: the global batch is world x batch_size, as with Lightning DDP).  `--checkpoint`
    resumes weights, Adam moments / step / learning rate, epoch and global step (Lightning's `fit(ckpt_path=...)`)."""
    distributed = int(os.environ.get('WORLD_SIZE', '1')) > 1
    if distributed and not dist.is_initialized():
        dist.init_process_group('nccl')
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(device)
    torch.manual_seed(8421)

    workers = int(getattr(args, 'num_workers', 0))
    sampler = _EpochSampler(len(train_data), rank, world, True)
    train_loader = DataLoader(train_data, args.batch_size, sampler=sampler, num_workers=workers, collate_fn=collate_fn,
                              pin_memory=True)
    val_loader = DataLoader(val_data, args.batch_size, False, num_workers=workers, collate_fn=collate_fn, pin_memory=True)
    model = model.to(device).train()
    scaler = getattr(model, 'loss_scaler', None)
    if scaler is not None and hasattr(scaler, 'set_batch_size'):
        scaler.set_batch_size(args.batch_size)      # the reference reads trainer.train_dataloader.batch_size
    trainer = FlatAdamTrainer(model)
    start_epoch, global_step = 0, 0
    if args.checkpoint:
        ckpt = torch.load(args.checkpoint, map_location=device)
        model.load_state_dict(ckpt['state_dict'])
        if 'trainer' in ckpt:
            trainer.load_state_dict(ckpt['trainer'])
        start_epoch = int(ckpt.get('epoch', 0))
        global_step = int(ckpt.get('global_step', 0))
        if scaler is not None and hasattr(scaler, 'set_global_step'):
            scaler.set_global_step(global_step)

    log_dir = Path(args.logs_dir) / 'lightning_logs' / (args.name or 'version_0')
    if rank == 0:
        log_dir.mkdir(parents=True, exist_ok=True)
        meta = {'Model type': args.model, 'N internal': args.n_internal, 'N boundary': args.n_boundary,
                'N observations': args.n_observations, 'Precision': args.precision, 'Batch size': args.batch_size}
        (log_dir / 'model_meta.json').write_text(json.dumps(meta, indent=4))

    def checkpoint(epoch_done: int) -> dict:
        return {'state_dict': model.state_dict(), 'trainer': trainer.state_dict(), 'epoch': epoch_done,
                'global_step': global_step}

    history = []
    graphed = bool(getattr(args, 'cuda_graph', True))
    copy_stream = torch.cuda.Stream(device=device)

    def upload(host_batch):
        """Host -> device copy of a collated batch on the copy stream (pinned by the DataLoader), one batch ahead of the
        step that consumes it."""
        with torch.cuda.stream(copy_stream):
            dev = model.transfer_batch_to_device(host_batch, device)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dev, ev

    epoch_seconds = getattr(args, 'epoch_seconds', None)      # optional list: wall time of every training epoch (bench.py)
    import time as _time
    # Optional: keep the whole training set in HBM and collate on the device (DeviceFoamDataset) --
    # `args.device_dataset = True`; with `args.geometry_cache = True` the batches also carry their cached FPS / ball-query
    # results (exact with this build's fixed FPS start).  The sample order is the same epoch_indices rule.
    device_data = None
    if getattr(args, 'device_dataset', False):
        from ..dataset.device_dataset import DeviceFoamDataset
        device_data = DeviceFoamDataset.from_samples([train_data[i] for i in range(len(train_data))], device=device)
        if getattr(args, 'geometry_cache', False) and model.executor.uses_geometry():
            device_data.build_geometry_cache(model)

    for epoch in range(start_epoch, args.epochs):
        model.train()
        sampler.set_epoch(epoch)
        res = None
        if epoch_seconds is not None:
            torch.cuda.synchronize()
            t_epoch = _time.perf_counter()
        if device_data is not None:
            order = epoch_indices(len(train_data), epoch, rank, world)
            for lo in range(0, len(order), args.batch_size):
                res = trainer.train_step(device_data.batch(order[lo:lo + args.batch_size]), graphed=graphed)
                global_step += 1
            it, pending = None, None
        else:
            it = iter(train_loader)
            nxt = next(it, None)
            pending = upload(nxt) if nxt is not None else None
        while pending is not None:
            batch, ev = pending
            nxt = next(it, None)
            pending = upload(nxt) if nxt is not None else None      # overlaps this step
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            batch.data.record_stream(cur)
            for v in batch.domain.values():
                v.record_stream(cur)
            res = trainer.train_step(batch, graphed=graphed)
            global_step += 1
        if epoch_seconds is not None:
            torch.cuda.synchronize()
            epoch_seconds.append(_time.perf_counter() - t_epoch)
        trainer.end_epoch()
        if rank == 0 and res is not None:
            history.append(float(res.loss))
        if val_loader is not None and len(val_data) > 0:
            model.eval()
            for batch in val_loader:
                model.validation_step(model.transfer_batch_to_device(batch, device))
        if rank == 0 and (epoch + 1) % 500 == 0:
            torch.save(checkpoint(epoch + 1), log_dir / f'checkpoint-{epoch}.ckpt')
    if rank == 0:
        torch.save(checkpoint(args.epochs), log_dir / 'model.ckpt')
    return history

"""Building blocks with the reference's constructor signatures and `state_dict` names
(reference models/modules.py), as PARAMETER CONTAINERS for the CUDA executor.

In the reference every block is an eager torch module; here a block only owns its parameters
(same names: 'Linear 0', 'Operator 0', 'Sa-0', 'Global-Sa', `lins.k`, ...) and knows how to
describe itself as a chain of jet layers (`engine.ChainLayer`).  The arithmetic is done by
libpcfd_sm100.so.  Calling a block directly runs its VALUE path through the same kernels (no
autograd graph) and raises off-GPU: there is no eager fallback.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch
from torch import Tensor, nn
from torch.nn import Dropout, Linear, Module, SiLU, Tanh

from .. import _lib, ops
from ..engine import ChainLayer, SALevel, SAStack, StepContext, chain_forward, mlp_chain, sa_forward
from ..ops import Jet


def activation_name(activation) -> Optional[str]:
    """Map an activation class / instance to the kernel's code name."""
    if activation is None:
        return None
    cls = activation if isinstance(activation, type) else type(activation)
    if issubclass(cls, SiLU):
        return 'silu'
    if issubclass(cls, Tanh):
        return 'tanh'
    raise NotImplementedError(f'activation {cls.__name__}: the jet kernels implement SiLU and Tanh '
                              '(the two the reference uses, models/pipn/pipn_foam.py:70, pipn_baseline.py:26)')


def get_batch(x: Tensor) -> Tensor:
    """PyG-style batch vector of a (B, N, *) tensor (reference models/modules.py:13-20)."""
    b, n = x.shape[0], x.shape[-2]
    return torch.arange(b, device=x.device, dtype=torch.int64).repeat_interleave(n)


def _value_forward(layers, pending_act, x: Tensor, pool: bool = False) -> Tensor:
    """Run a chain on plain features (..., K) through the kernels, value channel only.  With
    `pool`, x is (B, M, K) and the result is the max over M of the activated output, (B, 1, C)."""
    if not x.is_cuda:
        raise _lib.PcfdError('porous_cfd_b200 modules only execute on a CUDA sm_100 device (no CPU fallback)')
    lead, k = tuple(x.shape[:-1]), x.shape[-1]
    rows = 1
    for s in lead:
        rows *= s
    z0 = Jet.empty(1, rows, k, x.device)
    z0.t[0, :, :k].copy_(x.detach().reshape(rows, k))
    seg_len = lead[-1] if pool else 1
    zs = chain_forward(StepContext(x.device), layers, z0, seg_len)
    n = layers[-1].n
    if pool:
        vals, _ = ops.segmax_fwd(zs[-1].t[0], pending_act, None, rows // seg_len, seg_len, n)
        return vals[:, :n].reshape(*lead[:-1], 1, n)
    if pending_act is not None:
        vals, _ = ops.segmax_fwd(zs[-1].t[0], pending_act, None, rows, 1, n)   # length-1 segments: act only
        return vals[:, :n].reshape(*lead, n)
    return zs[-1].values().reshape(*lead, n)


class MLP(nn.Sequential):
    """Linear -> activation (-> Dropout) stack with the reference's child names
    ('Linear i', 'Activation i', 'Dropout i'; reference models/modules.py:23-53)."""

    def __init__(self, layers: list[int], dropout: list[float] = None, activation: type[Module] = Tanh,
                 last_activation=True):
        super().__init__()
        if dropout is not None and len(dropout) != len(layers) - 1:
            raise AssertionError(f'Mismatching number of layers ({len(layers)}) and dropout ({len(dropout)}).')
        self.sizes = list(layers)
        self.dropout_p = list(dropout) if dropout is not None else None
        self.act_name = activation_name(activation)
        self.last_activation = last_activation
        for i, (fan_in, fan_out) in enumerate(zip(layers[:-1], layers[1:])):
            self.add_module(f'Linear {i}', Linear(fan_in, fan_out))
            if last_activation or i + 2 < len(layers):
                self.add_module(f'Activation {i}', activation())
            if dropout is not None and dropout[i] > 0:
                self.add_module(f'Dropout {i}', Dropout(dropout[i]))

    def linears(self) -> list[Linear]:
        return [m for m in self.children() if isinstance(m, Linear)]

    def chain(self, first_act=None, first_drop=0.0):
        return mlp_chain(self.linears(), self.act_name, self.last_activation, self.dropout_p, first_act, first_drop)

    def forward(self, x: Tensor) -> Tensor:
        layers, (pend, _) = self.chain()
        return _value_forward(layers, pend, x)


class PointMLP(nn.Module):
    """Stand-in for torch_geometric.nn.MLP(channels, act=..., norm=None, plain_last=False) with the
    upstream parameter names `lins.k.{weight,bias}` (reference models/modules.py:506-512)."""

    def __init__(self, channel_list, act=None, norm=None, plain_last=False, dropout=0.0):
        super().__init__()
        if norm is not None or plain_last or (dropout and dropout > 0):
            raise NotImplementedError('only norm=None, plain_last=False, dropout=0 is used by the in-scope models')
        self.channel_list = list(channel_list)
        self.act_name = activation_name(act)
        self.lins = nn.ModuleList(Linear(i, o) for i, o in zip(channel_list[:-1], channel_list[1:]))

    def chain(self):
        return mlp_chain(list(self.lins), self.act_name, True)

    def forward(self, x: Tensor) -> Tensor:
        layers, (pend, _) = self.chain()
        return _value_forward(layers, pend, x)


class PointConvNext(nn.Module):
    """PointNetConv with radius-normalised relative positions and max aggregation
    (reference models/modules.py:277-292).  Executed by pcfd_sa_gather + jet GEMMs + pcfd_segmax."""

    def __init__(self, r: float, local_nn: PointMLP = None, **kwargs):
        super().__init__()
        self.r = r
        self.local_nn = local_nn


class SetAbstraction(nn.Module):
    """fps -> radius -> PointConvNext (reference models/modules.py:295-325)."""

    def __init__(self, ratio: float, r: float, mlp: PointMLP, max_neighbors=64):
        super().__init__()
        self.ratio, self.r, self.max_neighbors = ratio, r, max_neighbors
        self.conv = PointConvNext(r, local_nn=mlp)

    def level(self) -> SALevel:
        layers, (pend, _) = self.conv.local_nn.chain()
        return SALevel(self.ratio, self.r, layers, pend, self.max_neighbors)


class GlobalSetAbstraction(nn.Module):
    """MLP on [x, pos] then per-geometry max (reference models/modules.py:403-423)."""

    def __init__(self, mlp: PointMLP):
        super().__init__()
        self.nn = mlp


class SetAbstractionSeq(nn.Module):
    """'Sa-i' levels and an optional 'Global-Sa' (reference models/modules.py:483-527)."""

    def __init__(self, fraction: list[float], radius: list[float], conv_mlp: list[list[int]], return_skip=True,
                 activation: type[Module] = Tanh, max_neighbors=64):
        super().__init__()
        blocks = OrderedDict()
        for i, (frac, r, channels) in enumerate(zip(fraction, radius, conv_mlp)):
            blocks[f'Sa-{i}'] = SetAbstraction(frac, r, PointMLP(channels, act=activation()), max_neighbors)
        if len(conv_mlp) > len(radius):
            blocks['Global-Sa'] = GlobalSetAbstraction(PointMLP(conv_mlp[-1], act=activation()))
        self.layers = nn.Sequential(blocks)
        self.return_skip = return_skip
        self.act_name = activation_name(activation)

    def stack(self) -> SAStack:
        levels = [m.level() for m in self.layers if isinstance(m, SetAbstraction)]
        glob = [m for m in self.layers if isinstance(m, GlobalSetAbstraction)]
        global_layers = glob[0].nn.chain()[0] if glob else None
        return SAStack(levels, global_layers, self.act_name)


class BatchedDecorator(nn.Module):
    """(B, M, F) front end of a flattened-batch module (reference models/modules.py:85-98)."""

    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, x: Tensor, pos: Tensor) -> Tensor:
        """x (B, M, F), pos (B, M, D) -> (B, 1, E) pooled geometry feature (value path)."""
        if not x.is_cuda:
            raise _lib.PcfdError('porous_cfd_b200 modules only execute on a CUDA sm_100 device (no CPU fallback)')
        b, m, f = x.shape
        x0 = torch.empty((b * m, ops.round4(f)), dtype=torch.float32, device=x.device)
        x0[:, :f].copy_(x.detach().reshape(b * m, f))
        g, _ = sa_forward(StepContext(x.device), self.module.stack(), x0, x0.stride(0), f,
                          pos.detach().contiguous().float())
        e = self.module.stack().global_layers[-1].n
        return g[:, :e].reshape(b, 1, e)


class PointNetFeatureExtract(nn.Module):
    """Local shared MLP + global shared MLP + max-pool (reference models/modules.py:56-82)."""

    def __init__(self, local_layers: list[int], global_layers: list[int], activation: type[Module] = Tanh):
        super().__init__()
        self.local_feature = MLP(local_layers, activation=activation)
        self.global_feature = MLP(global_layers, activation=activation)


class PointNetFeatureExtractPp(nn.Module):
    """Local shared MLP + set-abstraction geometry encoder (reference models/modules.py:101-139)."""

    def __init__(self, local_layers: list[int], global_layers: list[list[int]], global_fraction: list[float],
                 global_radius: list[float], activation: type[Module] = Tanh, max_neighbors=64):
        super().__init__()
        self.local_feature = MLP(local_layers, activation=activation)
        self.global_feature = BatchedDecorator(SetAbstractionSeq(global_fraction, global_radius, global_layers,
                                                                 return_skip=False, activation=activation,
                                                                 max_neighbors=max_neighbors))


class GeometryEncoderPp(nn.Module):
    """PI-GANO++ geometry encoder (reference models/modules.py:142-168)."""

    def __init__(self, fraction: list[float], radius: list[float], conv_mlp: list[list[int]],
                 activation: type[Module] = Tanh, max_neighbors=64):
        super().__init__()
        self.set_abstraction = BatchedDecorator(SetAbstractionSeq(fraction, radius, conv_mlp, return_skip=False,
                                                                  activation=activation, max_neighbors=max_neighbors))

    def forward(self, x: Tensor, pos: Tensor) -> Tensor:
        return self.set_abstraction(x, pos)


class Branch(nn.Module):
    """Branch network: shared MLP + max over the parameter points (reference models/modules.py:171-190)."""

    def __init__(self, hidden_channels: list[int], activation: type[Module] = Tanh):
        super().__init__()
        self.linear = MLP(hidden_channels, activation=activation)

    def forward(self, param_features: Tensor) -> Tensor:
        layers, (pend, _) = self.linear.chain()
        return _value_forward(layers, pend, param_features, pool=True)


class GeometryEncoder(nn.Module):
    """PI-GANO geometry encoder: shared MLP on [x, pos] + max (reference models/modules.py:193-214)."""

    def __init__(self, hidden_channels: list[int], activation=Tanh):
        super().__init__()
        self.linear = MLP(hidden_channels, activation=activation)

    def forward(self, x: Tensor, pos: Tensor) -> Tensor:
        layers, (pend, _) = self.linear.chain()
        return _value_forward(layers, pend, torch.cat([x, pos], dim=-1), pool=True)


class NeuralOperator(nn.Module):
    """Dropout(act(Linear(h))) * branch embedding (reference models/modules.py:217-245)."""

    def __init__(self, out_channels: int, dropout: float, activation: type[Module] = Tanh):
        super().__init__()
        self.linear = nn.Sequential(nn.Linear(out_channels, out_channels))
        self.act_name = activation_name(activation)
        self.drop_p = float(dropout)
        if activation is not None:
            self.linear.append(activation())
        if dropout > 0:
            self.linear.append(nn.Dropout(dropout))


class NeuralOperatorSequential(nn.Sequential):
    """'Operator i' stack (reference models/modules.py:248-274)."""

    def __init__(self, n_operators: int, n_features: int, dropout: list[float], activation: type[Module] = Tanh,
                 last_activation=True):
        super().__init__()
        for i in range(n_operators):
            act = None if (i == n_operators - 1 and not last_activation) else activation
            self.add_module(f'Operator {i}', NeuralOperator(n_features, dropout[i], act))

    def operators(self) -> list[NeuralOperator]:
        return list(self.children())

"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference from /root/reference on CPU.

The reference (pure Python) depends on five packages that are not installed in this image and
cannot be installed (no network): lightning, torch_cluster, torch_geometric, foamlib, pyvista.
`install()` registers minimal stand-ins in `sys.modules` and puts the reference tree on
`sys.path`, after which `models.*`, `dataset.foam_data` and `dataset.foam_dataset` import and
`PorousPinnBase.training_step` (models/model_base.py:182-218) runs as written:

  * lightning.LightningModule -> torch.nn.Module + no-op `save_hyperparameters` / `log`
    (the training step only calls `self.log` through LossLogger, models/losses.py:139-146);
  * torch_cluster.{fps,radius}, torch_geometric.nn.{PointNetConv,MLP,global_max_pool,...},
    torch_geometric.utils.unbatch -> the CPU restatements in oracle/pyg_restate.py;
  * foamlib / pyvista -> empty placeholders (only the OpenFOAM parser touches them).

This only works where /root/reference exists (the build container).  It is used by
tests/golden/make_golden.py to produce the committed fixtures and by the `not gpu` tests that
re-check the stand-alone oracle (oracle/pinn_oracle.py) against the reference when it is present.
The GPU box has no /root/reference: nothing that runs there may call `install()`.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get('PCFD_REFERENCE_ROOT', '/root/reference')
_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'models'))


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install() -> None:
    """Make `import models...` resolve to the reference.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    import torch
    from torch import nn

    from oracle import pyg_restate as pr

    class LightningModule(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.logged = {}
            self.global_step = 0
            self.trainer = None
            self.logger = None

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, name, value, **k):
            self.logged[name] = value.detach().clone() if torch.is_tensor(value) else value

    class _Anything:
        def __init__(self, *a, **k):
            pass

    lightning = _module('lightning', LightningModule=LightningModule, Trainer=_Anything)
    pytorch = _module('lightning.pytorch')
    callbacks = _module('lightning.pytorch.callbacks', RichProgressBar=_Anything,
                        LearningRateMonitor=_Anything, ModelCheckpoint=_Anything)
    loggers = _module('lightning.pytorch.loggers', TensorBoardLogger=_Anything)
    lightning.pytorch = pytorch
    pytorch.callbacks, pytorch.loggers = callbacks, loggers

    _module('torch_cluster', fps=pr.fps, radius=pr.radius)
    tg = _module('torch_geometric')
    tg_nn = _module('torch_geometric.nn', PointNetConv=pr.PointNetConv, MLP=pr.PygMLP,
                    global_max_pool=pr.global_max_pool, knn_interpolate=pr.knn_interpolate,
                    Sequential=pr.PygSequential)
    tg_utils = _module('torch_geometric.utils', unbatch=pr.unbatch)
    tg.nn, tg.utils = tg_nn, tg_utils

    _module('foamlib', FoamCase=_Anything, FoamFile=_Anything)
    _module('pyvista', ArrayLike=object)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True

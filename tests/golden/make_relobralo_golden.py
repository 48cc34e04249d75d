"""Golden vectors of the reference's ReLoBRaLo scaler (models/losses.py:64-124), produced by running the UNMODIFIED
reference class (imported through oracle/ref_shim.py) in the build container:

    python tests/golden/make_relobralo_golden.py      -> tests/golden/relobralo.npz

Each case feeds a seeded sequence of positive loss vectors through RelobraloScaler.forward with the attributes the
reference reads from its LightningModule (global_step, trainer.train_dataloader.batch_size, logger.experiment,
training_loss_togger) provided by a stand-in, and stores the weighted vectors it returned.  beta is 1.0 or 0.0 so
that the torch.bernoulli draw is deterministic (rho = 1 / rho = 0).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pinn_oracle, ref_shim  # noqa: E402

CASES = {  # name: (num_losses, alpha, beta, tau, batch_size, steps, seed)
    'rho1_b1': (12, 0.95, 1.0, 1.0, 1, 9, 1),
    'rho0_b3': (9, 0.005, 0.0, 1.0, 3, 10, 2),
    'rho1_b2_tau': (6, 0.5, 1.0, 0.7, 2, 8, 3),
}


def main():
    ref_shim.install()
    from models.losses import RelobraloScaler
    out = {}
    for name, (n, alpha, beta, tau, bs, steps, seed) in CASES.items():
        g = torch.Generator().manual_seed(seed)
        seq = torch.rand(steps, n, generator=g) * torch.logspace(-3, 2, n) + 1e-4
        scaler = RelobraloScaler(n, alpha=alpha, beta=beta, tau=tau)
        model = types.SimpleNamespace(
            global_step=0,
            trainer=types.SimpleNamespace(train_dataloader=types.SimpleNamespace(batch_size=bs)),
            logger=types.SimpleNamespace(experiment=types.SimpleNamespace(add_scalars=lambda *a, **k: None)),
            training_loss_togger=types.SimpleNamespace(loss_labels=['Total loss'] + [f'l{i}' for i in range(n)]))
        got = []
        restated = pinn_oracle.Relobralo(n, alpha=alpha, rho=beta, tau=tau, batch_size=bs)
        for s in range(steps):
            model.global_step = s
            got.append(scaler(model, seq[s].clone()).detach().clone())
            mine = restated(seq[s])
            assert torch.allclose(mine, got[-1], rtol=1e-6, atol=0), (name, s, mine, got[-1])
        out[f'{name}/meta'] = np.array([n, alpha, beta, tau, bs, steps], dtype=np.float64)
        out[f'{name}/losses'] = seq.numpy()
        out[f'{name}/weighted'] = torch.stack(got).numpy()
    np.savez(os.path.join(ROOT, 'tests', 'golden', 'relobralo.npz'), **out)
    print('wrote relobralo.npz:', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()

"""A few eager, single-stream training steps of a BASELINE config (default: config 2, 32 geometries) -- the short command
that is run under ncu (launch list / --set full captures of selected kernels).
    python scripts/one_step.py [--config abc_pipn_pp] [--batch 32] [--steps 2]"""
import argparse
import os
import sys

os.environ.setdefault('PCFD_NO_OVERLAP', '1')
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import pcfd_import  # noqa: E402

pcfd_import.load()
import torch  # noqa: E402
from porous_cfd_b200 import factory, synthetic  # noqa: E402
from porous_cfd_b200.dataset.foam_data import FoamData  # noqa: E402

SHAPES = {'abc_pipn_pp': (1500, 1000, 700, 32), 'abc_pipn': (1500, 1000, 700, 13), 'duct_pigano': (1500, 1000, 700, 64),
          'windbreaks_pigano_pp': (16384, 8192, 4096, 2), 'manufactured_pipn_pp': (4096, 1024, 0, 32)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='abc_pipn_pp')
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--steps', type=int, default=2)
    args = ap.parse_args()
    ni, nb, no, b = SHAPES[args.config]
    b = args.batch or b
    spec = synthetic.model_spec(args.config)
    torch.manual_seed(3)
    model = factory.build_model(spec).cuda().train()
    data, labels, domain = synthetic.make_batch(spec['layout'], b, seed=0, n_internal=ni, n_boundary=nb, n_obs=no)
    batch = FoamData(data, labels, domain).to('cuda')
    for _ in range(args.steps):
        res = model.fused_step(batch)
    torch.cuda.synchronize()
    print('loss', float(res.loss))


if __name__ == '__main__':
    main()

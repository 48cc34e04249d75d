"""Development check of the vanilla-PIPN coupling path against the reference fixtures (run on the GPU box)."""
import sys, os, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import pcfd_import; pcfd_import.load()
from helpers import load_fixture, rel_l2, flat
from oracle import pinn_oracle
from porous_cfd_b200 import factory, synthetic
from porous_cfd_b200.dataset.foam_data import FoamData
for name in ['tiny_pipn', 'tiny_manufactured']:
    spec = synthetic.model_spec(name)
    data, domain, params, out = load_fixture(name)
    labels = synthetic.build_labels(spec['layout'])
    model = factory.build_model(spec); model.load_state_dict(params, strict=True); model = model.to('cuda').eval()
    batch = FoamData(data, labels, domain).to('cuda')
    ref = out['reference']
    for coup in (False, True):
        model.coupling = coup
        res = model.fused_step(batch, 'reference')
        torch.cuda.synchronize()
        rl = ((res.losses.cpu().double() - ref['losses'].double()).abs() / ref['losses'].double().abs())
        keys = list(params)
        grads = {k: model.executor.ctx.grads[id(p)].clone() for k, p in model.named_parameters()}
        print(name, 'coupling', coup, 'max loss rel', float(rl.max()), 'grad rel L2', rel_l2(flat(grads, keys), flat(ref['grads'], keys)))
        if coup:
            for k in keys:
                e = rel_l2(grads[k].double().cpu().flatten(), ref['grads'][k].double().flatten())
                print(f'     {k:50s} {e:.2e}  |ref| {float(ref["grads"][k].norm()):.3e}')

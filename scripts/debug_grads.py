import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import pcfd_import; pcfd_import.load()
import torch
from helpers import *
from porous_cfd_b200 import factory, synthetic
from porous_cfd_b200.dataset.foam_data import FoamData
names = sys.argv[1:] or ['tiny_pipn_pp']
for name in names:
  for mode in ['reference', 'true']:
    spec = synthetic.model_spec(name)
    data, domain, params, out = load_fixture(name)
    labels = synthetic.build_labels(spec['layout'])
    model = factory.build_model(spec); model.load_state_dict(params); model = model.to('cuda').eval()
    batch = FoamData(data, labels, domain).to('cuda')
    res = model.fused_step(batch, laplacian=mode)
    torch.cuda.synchronize()
    ref = out[mode]
    print(name, mode, 'loss terms rel err', ((res.losses.cpu()-ref['losses']).abs()/ref['losses'].abs()).max().item())
    tot = flat(ref['grads'], list(params)).norm().item()
    for k, p in model.named_parameters():
        g = model.executor.ctx.grads[id(p)].cpu().double(); r = ref['grads'][k].double()
        print(f'   {k:75s} |ref|={r.norm():.3e} err={(g-r).norm():.3e} rel={(g-r).norm()/(r.norm()+1e-30):.2e} share={(g-r).norm()/tot:.2e}')

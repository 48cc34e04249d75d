import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R,'tests'))
import pcfd_import; pcfd_import.load()
import torch
from helpers import *
from oracle import pinn_oracle
from porous_cfd_b200 import factory, synthetic, ops, engine
from porous_cfd_b200.ops import Jet
name='tiny_pipn_pp'
spec = synthetic.model_spec(name)
data, domain, params, out = load_fixture(name)
labels = synthetic.build_labels(spec['layout'])
model = factory.build_model(spec); model.load_state_dict(params); model = model.to('cuda').eval()
ex = model.executor
b=2; nb=24; d=3
bnd = pinn_oracle.rows(data, domain['boundary'])
bnd_c, bnd_id = pinn_oracle.field(bnd, labels,'C'), pinn_oracle.field(bnd, labels,'boundaryId')
geom = torch.cat([bnd_c,bnd_id],-1)
pl = {k:v.clone().requires_grad_(True) for k,v in params.items()}
tap=[]
import torch.nn.functional as F
g = pinn_oracle.set_abstraction_stack(geom, bnd_c, pl, 'feature_extract.global_feature.module.', {'radius':spec['fe_radius'],'fraction':spec['fe_fraction'],'layers':spec['fe_global_layers'],'max_neighbors':spec['max_neighbors']}, F.silu, tap)
gen = torch.Generator().manual_seed(0)
gg = torch.randn(g.shape, generator=gen)
(g*gg).sum().backward()
# cuda
stack = ex.plan['sa_stack']
x0 = torch.zeros(b*nb, 8, device='cuda'); x0[:, :7] = geom.reshape(-1,7).cuda()
ex.flat_grad.zero_()
gd, saved = engine.sa_forward(ex.ctx, stack, x0, 8, 7, bnd_c.contiguous().cuda())
print('fwd err', (gd[:, :32].cpu()-g[:,0]).abs().max().item())
saved['debug']={}
ggd = gg[:,0].contiguous().cuda()
engine.sa_backward(ex.ctx, stack, saved, ggd, 32)
torch.cuda.synchronize()
gprev, gein = saved['debug'][1]
print('x1 grad: ref norm', tap[0].grad.norm().item(), 'err', (gprev[:, :24].cpu()-tap[0].grad).norm().item())
print((gprev[:, :24].cpu()-tap[0].grad).abs().max(dim=1))
print('slots l1', saved['levels'][1]['slots'])
print('arg l1', saved['levels'][1]['arg'][:, :8])
for k,p in model.named_parameters():
    if 'Sa-' in k or 'Global' in k:
        gr = pl[k].grad; gc = ex.ctx.grads[id(p)].cpu()
        print(k, (gc-gr).norm().item()/gr.norm().item())

# ---- isolate level 0: recompute its backward with torch autograd on the GPU from the saved tensors
import torch.nn.functional as F
sv = saved['levels'][0]
ein = sv['zs'][0].t[0, :, :10].clone()
slots = sv['slots']
W0 = params['feature_extract.global_feature.module.layers.Sa-0.conv.local_nn.lins.0.weight'].cuda().requires_grad_(True)
b0 = params['feature_extract.global_feature.module.layers.Sa-0.conv.local_nn.lins.0.bias'].cuda().requires_grad_(True)
W1 = params['feature_extract.global_feature.module.layers.Sa-0.conv.local_nn.lins.1.weight'].cuda().requires_grad_(True)
b1 = params['feature_extract.global_feature.module.layers.Sa-0.conv.local_nn.lins.1.bias'].cuda().requires_grad_(True)
h = F.silu(F.linear(F.silu(F.linear(ein, W0, b0)), W1, b1)).reshape(24, 9, 24)
h = torch.where((slots >= 0)[:, :, None], h, torch.full_like(h, -float('inf')))
mx, am = h.max(dim=1)
print('arg equal', torch.equal(am.int(), sv['arg']), 'x1 equal', (mx - saved['levels'][1]['zs'][0].t.new_tensor(0)).shape)
(mx * gprev[:, :24]).sum().backward()
names = ['lins.0.weight','lins.0.bias','lins.1.weight','lins.1.bias']
for nm, t in zip(names, [W0,b0,W1,b1]):
    k = 'feature_extract.global_feature.module.layers.Sa-0.conv.local_nn.'+nm
    p = dict(model.named_parameters())[k]
    gc = ex.ctx.grads[id(p)]
    print(nm, 'kernel vs gpu-autograd', ((gc - t.grad).norm()/t.grad.norm()).item(), ' gpu-autograd vs oracle', ((t.grad.cpu()-pl[k].grad).norm()/pl[k].grad.norm()).item())
# compare max values between oracle tap and kernel
print('x1 value err', (mx.cpu() - tap[0].detach()).abs().max().item())
# tie check in oracle: count of slots attaining max
cnt = (h == mx[:, None, :]).sum(dim=1)
print('ties:', (cnt > 1).sum().item(), 'of', cnt.numel())

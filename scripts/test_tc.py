"""Quick check of the tcgen05 engine against the FFMA engine (run on the GPU box)."""
import sys, os, math
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import pcfd_import; pcfd_import.load()
import torch
from porous_cfd_b200 import ops
from porous_cfd_b200.ops import Jet

which = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
CASES = [(1, 1024, 64, 128, None), (4, 1500, 64, 384, 'silu'), (4, 3000, 384, 128, 'silu'), (1, 2000, 131, 256, 'silu'),
         (3, 640, 176, 352, 'silu'), (7, 700, 64, 128, 'tanh'), (5, 512, 40, 24, 'tanh'), (1, 8500, 10, 64, None),
         (4, 48000, 128, 4, 'silu'), (1, 4000, 259, 1024, 'silu')]
torch.manual_seed(0)
for cj, rows, k, n, act in CASES:
    rpg = rows // 2 if rows % 2 == 0 else 0
    zin = Jet.empty(cj, rows, k, 'cuda'); zin.t.normal_()
    w = torch.randn(n, k + 8, device='cuda') / math.sqrt(k)
    bias = torch.randn(n, device='cuda')
    esc = torch.randn(2, k, device='cuda') if rpg else None
    tin = ops.make_intrans(act, 0, esc) if (act or esc is not None) else None
    res = {}
    for eng in (0, 1):
        ops.set_gemm_engine(eng)
        if which == 'fwd':
            out = ops.jet_linear_fwd(zin, tin, w, 4, k, bias, None, rpg, n)
            res[eng] = out.t[:, :, :n].double().cpu()
        torch.cuda.synchronize()
    ref = res[0]
    err = (res[1] - ref).norm() / ref.norm()
    print(f'{which} cj={cj} rows={rows} k={k} n={n} act={act}: rel err tc vs ffma = {err:.3e}  maxabs {float((res[1]-ref).abs().max()):.3e}')
ops.set_gemm_engine(0)

"""Quick check of the tcgen05 engine against the FFMA engine (run on the GPU box)."""
import sys, os, math
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import pcfd_import; pcfd_import.load()
import torch
from porous_cfd_b200 import ops
from porous_cfd_b200.ops import Jet

CASES = [(1, 1024, 64, 128, None), (4, 1500, 64, 384, 'silu'), (4, 3000, 384, 128, 'silu'), (1, 2000, 131, 256, 'silu'),
         (3, 640, 176, 352, 'silu'), (7, 700, 64, 128, 'tanh'), (5, 512, 40, 24, 'tanh'), (5, 1100, 176, 176, 'silu'),
         (1, 8500, 64, 128, 'silu'), (4, 48000, 128, 4, 'silu'), (1, 4000, 259, 1024, 'silu'), (7, 3000, 512, 512, 'silu'),
         (4, 2048, 64, 64, 'silu'), (1, 1537, 1024, 384, None)]
torch.manual_seed(0)
worst = 0.0
for cj, rows, k, n, act in CASES:
    rpg = rows // 2 if rows % 2 == 0 else 0
    ng = 2 if rpg else 1
    zin = Jet.empty(cj, rows, k, 'cuda'); zin.t.normal_()
    gz = Jet.empty(cj, rows, n, 'cuda'); gz.t.normal_()
    w = torch.randn(n, k + 8, device='cuda') / math.sqrt(k)
    bias = torch.randn(n, device='cuda')
    esc = torch.randn(ng, k, device='cuda') if rpg else None
    tin = ops.make_intrans(act, 0, esc) if (act or esc is not None) else None
    res = {}
    for eng in (0, 1):
        ops.set_gemm_engine(eng)
        out = ops.jet_linear_fwd(zin, tin, w, 4, k, bias, None, rpg, n)
        ge = torch.zeros(ng, k, device='cuda') if esc is not None else None
        gzin = ops.jet_linear_bwd_dx(gz, w, 4, zin, tin, ge, rpg, k, n)
        gw = torch.zeros_like(w); gb = torch.zeros(n, device='cuda')
        ws = torch.empty(ops.dw_workspace_bytes(cj, rows, rpg, k, n), dtype=torch.uint8, device='cuda')
        ops.jet_linear_bwd_dw(gz, zin, tin, gw, 4, gb, None, rpg, k, n, ws)
        torch.cuda.synchronize()
        res[eng] = [out.t[:, :, :n].double().cpu(), gzin.t[:, :, :k].double().cpu(), gw.double().cpu(), gb.double().cpu(),
                    ge.double().cpu() if ge is not None else torch.zeros(1, dtype=torch.float64)]
    errs = [float((a - b).norm() / (b.norm() + 1e-30)) for a, b in zip(res[1], res[0])]
    worst = max(worst, max(errs))
    print(f'cj={cj} rows={rows} k={k} n={n} act={act}: fwd {errs[0]:.2e} dx {errs[1]:.2e} dw {errs[2]:.2e} db {errs[3]:.2e} de {errs[4]:.2e}')
ops.set_gemm_engine(0)
print('WORST', worst)

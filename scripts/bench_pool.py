"""Timing + error of the sparse (last layer -> max pool) backward against the dense form it replaces, on the pooled
layers of the BASELINE configs (run on the GPU box):  python scripts/bench_pool.py"""
import math
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import pcfd_import  # noqa: E402

pcfd_import.load()
import torch  # noqa: E402
from porous_cfd_b200 import ops  # noqa: E402
from porous_cfd_b200.ops import Jet  # noqa: E402

CASES = [  # name, n_seg, seg_len, k, c
    ('abc sa0 L1', 16000, 17, 64, 128),
    ('abc sa1 L1', 4000, 17, 128, 256),
    ('abc glob L1', 32, 125, 256, 1024),
    ('windbreaks sa0 L1 (B=2)', 8192, 65, 64, 128),
    ('manufactured sa0 (B=32, 1024 bnd)', 16384, 65, 8, 64),
]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    torch.manual_seed(0)
    for name, n_seg, L, k, c in CASES:
        rows = n_seg * L
        zj = Jet.empty(1, rows, k, 'cuda'); zj.t.normal_()
        w = torch.randn(c, k, device='cuda') / math.sqrt(k)
        b = torch.randn(c, device='cuda')
        tin = ops.make_intrans('silu')
        zl = ops.jet_linear_fwd(zj, tin, w, 0, k, b, None, 0, c)
        pooled, arg, zsel = ops.segmax_fwd_z(zl.t[0], 'silu', None, n_seg, L, c)
        gout = torch.randn(n_seg, c, device='cuda')
        gw, gb = torch.zeros_like(w), torch.zeros(c, device='cuda')
        ws = torch.empty(max(ops.pool_layer_bwd_workspace_bytes(n_seg, L, k, c), ops.dw_workspace_bytes(1, rows, 0, k, c)),
                         dtype=torch.uint8, device='cuda')
        t_dw = timeit(lambda: ops.pool_layer_bwd(gout, gout.stride(0), arg, zsel, 'silu', n_seg, L, c, zj, tin, k, w, gw, gb, False, ws))
        t_dx = timeit(lambda: ops.pool_layer_bwd(gout, gout.stride(0), arg, zsel, 'silu', n_seg, L, c, zj, tin, k, w, None, None, True, None))
        t_sb = timeit(lambda: ops.segmax_bwd(gout, gout.stride(0), arg, zl.t[0], 'silu', n_seg, L, c))
        gz = Jet(ops.segmax_bwd(gout, gout.stride(0), arg, zl.t[0], 'silu', n_seg, L, c), c)
        t_ddx = timeit(lambda: ops.jet_linear_bwd_dx(gz, w, 0, zj, tin, None, 0, k, c))
        t_ddw = timeit(lambda: ops.jet_linear_bwd_dw(gz, zj, tin, gw, 0, gb, None, 0, k, c, ws))
        # errors of the sparse form against the dense one
        gw1, gb1 = torch.zeros_like(w), torch.zeros(c, device='cuda')
        gzs = ops.pool_layer_bwd(gout, gout.stride(0), arg, zsel, 'silu', n_seg, L, c, zj, tin, k, w, gw1, gb1, True, ws)
        gw2, gb2 = torch.zeros_like(w), torch.zeros(c, device='cuda')
        ops.jet_linear_bwd_dw(gz, zj, tin, gw2, 0, gb2, None, 0, k, c, ws)
        gzd = ops.jet_linear_bwd_dx(gz, w, 0, zj, tin, None, 0, k, c)
        e = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm())
        per_block = [(i, e(gw1[i:i + 64], gw2[i:i + 64])) for i in range(0, c, 64)]
        print(f'{name:36s} sparse dw {t_dw:7.1f} us  dx {t_dx:7.1f} us | dense segmax_bwd {t_sb:7.1f} dx {t_ddx:7.1f} dw {t_ddw:7.1f} us'
              f' | err gw {e(gw1, gw2):.1e} gb {e(gb1, gb2):.1e} gzin {e(gzs.t[0, :, :k], gzd.t[0, :, :k]):.1e}')
        bad = [(i, round(x, 4)) for i, x in per_block if x > 1e-4]
        if bad:
            print('   channel blocks off:', bad[:20])


if __name__ == '__main__':
    main()

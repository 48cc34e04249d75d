"""Per-layer check + timing of the jet GEMM engines on the layer shapes of the BASELINE configs (run on the
GPU box):  python scripts/bench_layers.py [--engines 0,2] [--passes fwd,dx,dw] [--set abc|all] [--iters 20]

For every (cj, rows, k, n, act, dropout) it runs each pass on each engine, reports the relative L2 error
against engine 0 (fp32 FFMA) and the CUDA-event time, algorithmic TFLOP/s and HBM GB/s."""
import argparse
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

# (name, cj, rows, k, n, act, drop_p)
ABC = [
    ('int L1 64->64', 4, 48000, 64, 64, 'silu', 0.0),
    ('int L2 64->384', 4, 48000, 64, 384, 'silu', 0.0),
    ('int L3 384->128', 4, 48000, 384, 128, 'silu', 0.03),
    ('bnd L2 64->384', 1, 32000, 64, 384, 'silu', 0.0),
    ('bnd L3 384->128', 1, 32000, 384, 128, 'silu', 0.03),
    ('sa0 L0 10->64', 1, 272000, 10, 64, None, 0.0),
    ('sa0 L1 64->128', 1, 272000, 64, 128, 'silu', 0.0),
    ('sa1 L0 131->128', 1, 68000, 131, 128, None, 0.0),
    ('sa1 L1 128->256', 1, 68000, 128, 256, 'silu', 0.0),
    ('glob L0 259->256', 1, 4000, 259, 256, None, 0.0),
    ('glob L1 256->1024', 1, 4000, 256, 1024, 'silu', 0.0),
    ('int first 3->64', 4, 48000, 3, 64, None, 0.0),
    ('int last 128->4', 4, 48000, 128, 4, 'silu', 0.0),
    ('bnd last 128->4', 1, 32000, 128, 4, 'silu', 0.0),
    ('cvec 1024->384', 1, 32, 1024, 384, None, 0.0),
]
OTHERS = [
    ('pigano 176->176', 3, 96000, 176, 176, 'silu', 0.0),
    ('pigano 352->352', 3, 96000, 352, 352, 'silu', 0.1),
    ('windbreaks 512->512 true', 7, 16384, 512, 512, 'silu', 0.15),
    ('manuf 512->256 true', 5, 65536, 512, 256, 'tanh', 0.0),
    ('odd rows', 4, 1511, 64, 384, 'silu', 0.0),
    ('odd n', 1, 5000, 96, 200, 'silu', 0.0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--engines', default="0,2")
    ap.add_argument('--passes', default='fwd,dx,dw')
    ap.add_argument('--set', default='abc')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--escale', type=int, default=0)
    ap.add_argument('--only', default='')
    ap.add_argument('--nograph', type=int, default=0)
    args = ap.parse_args()
    engines = [int(e) for e in args.engines.split(',')]
    passes = args.passes.split(',')
    cases = ABC + (OTHERS if args.set == 'all' else [])
    if args.only:
        cases = [c for c in ABC + OTHERS if args.only in c[0]]
    torch.manual_seed(0)
    seed_dev = torch.full((1,), 1234, dtype=torch.int64, device='cuda')
    tot = {e: {p: 0.0 for p in passes} for e in engines}
    worst = 0.0
    for name, cj, rows, k, n, act, drop in cases:
        # 32 geometries per batch as in the bench workload (per-geometry bias / cvec sums scale with it)
        ng = 32 if rows % 32 == 0 else (2 if rows % 2 == 0 else 1)
        rpg = rows // ng if ng > 1 else 0
        zin = Jet.empty(cj, rows, k, 'cuda')
        zin.t.normal_()
        gz = Jet.empty(cj, rows, n, 'cuda')
        gz.t.normal_()
        w = torch.randn(n, k, device='cuda') / math.sqrt(k)
        if k % 4:
            wp = torch.zeros(n, ops.round4(k), device='cuda')
            wp[:, :k] = w
            w = wp
        bias = torch.randn(n, device='cuda')
        esc = torch.randn(ng, k, device='cuda') if (args.escale and rpg) else None
        tin = ops.make_intrans(act, 0, esc, drop, seed_dev, 7) if (act or esc is not None or drop > 0) else None
        ws = torch.empty(ops.dw_workspace_bytes(cj, rows, rpg, k, n), dtype=torch.uint8, device='cuda')
        flops = 2.0 * cj * rows * k * n
        byts = {'fwd': 4.0 * cj * rows * (k + n), 'dx': 4.0 * cj * rows * (n + 2 * k), 'dw': 4.0 * cj * rows * (k + n)}
        ref = {}
        line = f'{name:26s} cj={cj} rows={rows:6d} {k:4d}->{n:4d}'
        print(line)
        for eng in engines:
            ops.set_gemm_engine(eng)
            for p in passes:
                def run():
                    if p == 'fwd':
                        return ops.jet_linear_fwd(zin, tin, w, 0, k, bias, None, rpg, n).t[:, :, :n]
                    if p == 'dx':
                        ge = torch.zeros(ng, k, device='cuda') if esc is not None else None
                        return ops.jet_linear_bwd_dx(gz, w, 0, zin, tin, ge, rpg, k, n).t[:, :, :k]
                    gw = torch.zeros_like(w)
                    gb = torch.zeros(n, device='cuda')
                    ops.jet_linear_bwd_dw(gz, zin, tin, gw, 0, gb, None, rpg, k, n, ws)
                    return torch.cat([gw[:, :k].flatten(), gb])
                try:
                    out = run().double()
                    torch.cuda.synchronize()
                except Exception as exc:  # noqa: BLE001
                    print(f'    engine {eng} {p}: FAILED {exc}')
                    continue
                if eng == engines[0]:
                    ref[p] = out
                    err = 0.0
                else:
                    err = float((out - ref[p]).norm() / (ref[p].norm() + 1e-30))
                    worst = max(worst, err)
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if args.nograph:
                    a.record()
                    for _ in range(args.iters):
                        run()
                    b.record()
                    torch.cuda.synchronize()
                else:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for _ in range(args.iters):
                            run()
                    g.replay()
                    torch.cuda.synchronize()
                    a.record()
                    g.replay()
                    b.record()
                    torch.cuda.synchronize()
                    del g
                us = a.elapsed_time(b) * 1e3 / args.iters
                tot[eng][p] += us
                print(f'    engine {eng} {p:3s}: {us:9.1f} us  {flops / us * 1e-6:7.1f} TFLOP/s  {byts[p] / us * 1e-3:7.0f} GB/s  '
                      f'rel err vs e{engines[0]} {err:.2e}')
        del zin, gz, ws
    ops.set_gemm_engine(0)
    print('TOTAL us per pass:', {e: {p: round(v, 1) for p, v in d.items()} for e, d in tot.items()})
    print('WORST', worst)


if __name__ == '__main__':
    main()

"""FPS timing on the point-set sizes of the BASELINE configs: python scripts/bench_fps.py  (GPU box).
PCFD_FPS_BUCKET_MIN=0 gives the plain one-CTA-per-geometry kernels for comparison."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import pcfd_import  # noqa: E402

pcfd_import.load()
import torch  # noqa: E402
from porous_cfd_b200 import ops  # noqa: E402

CASES = [(32, 1000, 3, 0.5), (32, 500, 3, 0.25), (2, 8192, 3, 0.5), (2, 4096, 3, 0.25), (8, 8192, 3, 0.5), (32, 1024, 2, 0.5),
         (32, 4096, 2, 0.5), (4, 16384, 2, 0.5), (32, 16384, 2, 0.5), (8, 65536, 2, 0.5), (32, 2500, 3, 0.5)]


def main():
    for nb, n, d, ratio in CASES:
        g = torch.Generator().manual_seed(n)
        pos = (torch.rand(nb, n, d, generator=g) * 2 - 1).cuda()
        for _ in range(2):
            ops.fps(pos, ratio)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 5
        a.record()
        for _ in range(it):
            ops.fps(pos, ratio)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / it
        m = int(-(-ratio * n // 1))
        print(f'fps  B={nb:3d} n={n:6d} d={d} -> m={m:6d}: {ms:8.3f} ms  ({ms * 1e3 / m:6.3f} us / sample)')


if __name__ == '__main__':
    main()

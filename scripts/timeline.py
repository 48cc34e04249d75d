"""Kernel timeline of the captured step graph (torch.profiler / CUPTI activity records of graph replays): every kernel with
its start, duration and stream, so that gaps, overlap between the two streams and the critical path can be read off.
Run on the GPU box:   python scripts/timeline.py [--config abc_pipn_pp] [--batch 32] > gpurun_out/timeline.csv"""
import argparse
import os
import sys

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
    ap.add_argument('--eval', action='store_true')
    ap.add_argument('--pipeline', type=int, default=1)
    ap.add_argument('--dp', action='store_true', help='under torchrun: all-reduce + Adam behind every step (rank 0 prints)')
    args = ap.parse_args()
    ni, nb, no, b = SHAPES[args.config]
    b = args.batch or b
    spec = synthetic.model_spec(args.config)
    torch.manual_seed(3)
    rank = 0
    if args.dp:
        import torch.distributed as dist
        rank = int(os.environ.get('RANK', '0'))
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
    model = factory.build_model(spec).cuda()
    model = model.eval() if args.eval else model.train()
    ex = model.executor
    model.pipeline_geometry = bool(args.pipeline) and ex.uses_geometry()
    trainer = None
    if args.dp:
        from porous_cfd_b200.common.training import FlatAdamTrainer
        trainer = FlatAdamTrainer(model)
    batches = []
    for i in range(2):
        data, labels, domain = synthetic.make_batch(spec['layout'], b, seed=i + 10 * rank, n_internal=ni, n_boundary=nb, n_obs=no)
        batches.append(FoamData(data, labels, domain).to('cuda'))
    def tail():
        if trainer is not None:
            trainer.reduce_gradients()
            trainer.step()

    for i in range(6):
        cur, nxt = batches[i % 2], batches[(i + 1) % 2]
        ex.graphed_step(cur.data, cur.labels, cur.domain, 'reference', next_batch=nxt)
        tail()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            cur, nxt = batches[i % 2], batches[(i + 1) % 2]
            ex.graphed_step(cur.data, cur.labels, cur.domain, 'reference', next_batch=nxt)
            tail()
        torch.cuda.synchronize()
    if rank != 0:
        return
    import json
    out_dir = os.path.join(R, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f'trace_{args.config}_{os.getpid()}.json')
    prof.export_chrome_trace(path)
    tr = json.load(open(path))
    evs = [e for e in tr['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and e.get('ph') == 'X']
    evs.sort(key=lambda e: e['ts'])
    t0 = evs[0]['ts'] if evs else 0
    print('start_us,dur_us,stream,name')
    for e in evs:
        name = e['name'].replace(',', ';')
        print(f"{e['ts'] - t0:.2f},{e['dur']:.2f},{e.get('args', {}).get('stream', -1)},{name[:160]}")
    os.remove(path)


if __name__ == '__main__':
    main()

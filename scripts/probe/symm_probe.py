"""Probe (torchrun, N >= 2): does torch's symmetric memory work on this box?  Allocates a symmetric buffer, exchanges
handles, writes into the peer's copy and reads it back; prints the pointer table and multicast support."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank = int(os.environ['RANK'])
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    world = dist.get_world_size()
    group = dist.group.WORLD
    try:
        symm_mem.enable_symm_mem_for_group(group.group_name)
    except Exception as e:
        print(rank, 'enable_symm_mem_for_group:', type(e).__name__, e)
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    hdl = symm_mem.rendezvous(t, group.group_name)
    print(rank, 'rank/world', hdl.rank, hdl.world_size, 'buffer_ptrs', [hex(p) for p in hdl.buffer_ptrs],
          'multicast', hex(hdl.multicast_ptr) if hdl.multicast_ptr else 0, 'signal_pad_size', hdl.signal_pad_size)
    try:
        print(rank, 'has_multicast_support', symm_mem._SymmetricMemory.has_multicast_support(symm_mem.DeviceType.CUDA, local))
    except Exception as e:
        print(rank, 'has_multicast_support probe:', type(e).__name__, e)
    hdl.barrier()
    peer = (rank + 1) % world
    pb = hdl.get_buffer(peer, (1 << 20,), torch.float32)
    got = float(pb[:16].sum()) / 16
    print(rank, 'peer value read over P2P', got, 'expected', peer + 1)
    hdl.barrier()
    pb[rank * 16:(rank + 1) * 16] = 100.0 + rank          # write into the peer's buffer
    hdl.barrier()
    torch.cuda.synchronize()
    src = (rank - 1) % world
    print(rank, 'value written by peer', float(t[src * 16]), 'expected', 100.0 + src)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    sys.exit(main())

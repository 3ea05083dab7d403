"""Per-rank pinned host <-> device bandwidth with 1 / 2 / 4 / 8 ranks copying at once (torchrun, one rank per GPU).
Explains the end-to-end (host-buffer) scaling of bench.py: the per-rank rate when all ranks of the box copy is the
ceiling of `e2e`.  Usage: python -m torch.distributed.run --nproc-per-node 8 tools/pcie_probe_ranks.py [--no-bind]"""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'flash-fft-conv_b200')]
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
bind = '--no-bind' not in sys.argv
note = 'unbound'
if bind:
    from flashfftconv.parallel import bind_to_gpu_numa_node
    note = bind_to_gpu_numa_node(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = 226492416                       # the C2 input of one step
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1); h_out.fill_(1)       # first touch on this rank's node
d_in = torch.empty(n, dtype=torch.uint8, device='cuda'); d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d(): d_in.copy_(h_in, non_blocking=True)
def d2h(): h_out.copy_(d_out, non_blocking=True)
def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


def timed(fn, active, reps=6):
    if active:
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if active:
        for _ in range(reps):
            fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps if active else 0.0], device='cuda')
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


if rank == 0:
    print(f'# {world} ranks, {"NUMA-bound" if bind else "unbound"} ({note}); {n / 1e6:.0f} MB per copy; GB/s per rank (slowest rank)')
    print('| ranks copying at once | H2D | D2H | both directions (per direction) |\n|---|---|---|---|')
levels = [c for c in (1, 2, 4, 8) if c <= world]
for c in levels:
    # spread the active ranks over both sockets the way torchrun fills the box: ranks 0..c-1
    active = rank < c
    r = [timed(fn, active) for fn in (h2d, d2h, both)]
    if rank == 0:
        print(f'| {c} | {n / r[0] / 1e6:.1f} | {n / r[1] / 1e6:.1f} | {n / r[2] / 1e6:.1f} |', flush=True)
dist.destroy_process_group()

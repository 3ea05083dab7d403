"""Why is the training-mode C2 step slower under torchrun (0.18-0.20 ms) than alone (0.158 ms)?  Times the step with CUDA
events and wall clock, with and without process-group / NUMA binding, at several loop lengths."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'flash-fft-conv_b200')]
import __graft_entry__ as ge
ge.build()
from flashfftconv import FlashFFTConv
rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
mode = sys.argv[1] if len(sys.argv) > 1 else 'plain'
if 'bind' in mode:
    from flashfftconv.parallel import bind_to_gpu_numa_node
    print(rank, bind_to_gpu_numa_node(local), 'affinity', len(os.sched_getaffinity(0)), flush=True)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dist.barrier()
N, B, H = 8192, 16, 768
conv = FlashFFTConv(N, dtype=torch.bfloat16).cuda()
u = torch.randn(B, H, N, device='cuda').to(torch.bfloat16); k = torch.randn(H, N, device='cuda') / N ** 0.5


def run(n, training):
    conv.train(training)
    for _ in range(5):
        conv(u, k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        conv(u, k)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, (t1 - t0) / n * 1e6


for n in (10, 20, 100, 400):
    a = run(n, True); b = run(n, False)
    print(f'rank {rank}/{world} [{mode}] steps={n}: train {a[0]:.1f} us/step (host enqueue {a[1]:.1f} us/step), eval {b[0]:.1f} us/step (host {b[1]:.1f})', flush=True)
if world > 1:
    dist.destroy_process_group()

"""CPU cost of enqueueing one module call (tiny shapes: the GPU is never the bottleneck, so wall time per call = host
overhead of the Python mirror + C ABI + launches).  Compare with the GPU time of a step (C2: 0.16 ms)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'flash-fft-conv_b200')]
import __graft_entry__ as ge
ge.build()
from flashfftconv import FlashFFTConv
import cProfile, pstats

for N, gated in ((8192, False), (8192, True), (1024, True), (32768, False)):
    conv = FlashFFTConv(N, dtype=torch.bfloat16).cuda()
    u = torch.randn(2, 2, N, device='cuda').to(torch.bfloat16)
    k = torch.randn(2, N, device='cuda')
    g = [torch.randn(2, 2, N, device='cuda').to(torch.bfloat16) for _ in range(2)] if gated else []
    for mode in ('train', 'eval'):
        conv.train(mode == 'train')
        for _ in range(50):
            conv(u, k, *g)
        torch.cuda.synchronize()
        n = 2000
        t = time.perf_counter()
        for _ in range(n):
            conv(u, k, *g)
        dt = time.perf_counter() - t
        torch.cuda.synchronize()
        print(f'N={N} gated={gated} {mode}: {dt / n * 1e6:.1f} us host time per forward call', flush=True)
conv = FlashFFTConv(8192, dtype=torch.bfloat16).cuda()
u = torch.randn(2, 2, 8192, device='cuda').to(torch.bfloat16); k = torch.randn(2, 8192, device='cuda')
pr = cProfile.Profile(); pr.enable()
for _ in range(2000):
    conv(u, k)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)

"""Where does a module-level step go?  Times (CUDA events, L2 flushed by rotating inputs) for the C2 workload:
kernel only, rfft only, kf pack only, full module call, and the module call replayed from a CUDA graph."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
from flashfftconv import FlashFFTConv, _lib
from flashfftconv.conv import _pack_kf, _ptr, _stream
N = int(os.environ.get('N', 8192)); B = int(os.environ.get('B', 16)); H = int(os.environ.get('H', 768)); L = N
dev = torch.device('cuda')
mod = FlashFFTConv(N, dtype=torch.bfloat16).cuda(); plan = mod.plan(dev)
us = [torch.randn(B, H, L, device=dev).to(torch.bfloat16) for _ in range(3)]
k = torch.randn(H, L, device=dev) / L ** 0.5
kf = _pack_kf(mod, plan, k, 0); y = torch.empty_like(us[0])
def timeit(fn, n=30):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
def kern(i): _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(us[i % 3]), _ptr(kf), None, None, _ptr(y), B, H, L, None, 0, _stream()))
print('kernel only   %.1f us' % timeit(kern))
print('rfft only     %.1f us' % timeit(lambda i: torch.fft.rfft(k, n=N)))
print('kf pack(+rfft) %.1f us' % timeit(lambda i: _pack_kf(mod, plan, k, 0)))
print('module call   %.1f us' % timeit(lambda i: mod(us[i % 3], k)))
import time
torch.cuda.synchronize(); t = time.perf_counter()
for i in range(200): mod(us[i % 3], k)
t_host = (time.perf_counter() - t) / 200 * 1e6
torch.cuda.synchronize()
print('module call host-side issue time %.1f us (async)' % t_host)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(3): mod(us[0], k)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        yg = mod(us[0], k)
torch.cuda.synchronize()
print('graph replay  %.1f us' % timeit(lambda i: g.replay()))

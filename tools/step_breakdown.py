"""Where does a module-level step go?  CUDA-event times for the C2 workload (inputs rotate through 3 buffers > L2):
kernel only, the filter-side launches, the full module call forward and forward+backward."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
from flashfftconv import FlashFFTConv, _lib
from flashfftconv.conv import _pack_kf, _pack_kf_from_natural, _kf_natural, _ptr, _stream
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
print('kernel only                 %.1f us' % timeit(kern))
print('rfft only                   %.1f us' % timeit(lambda i: torch.fft.rfft(k, n=N)))
print('rfft + bffc_kf_pack_rfft    %.1f us' % timeit(lambda i: _pack_kf_from_natural(mod, plan, _kf_natural(mod, k), 0)))
print('k -> k_f (library path)     %.1f us' % timeit(lambda i: _pack_kf(mod, plan, k, 0)))
if N <= 8192:
    kf_out = torch.empty((H, 8192), dtype=torch.int32, device=dev)
    print('bffc_kf_from_filter (direct) %.1f us' % timeit(lambda i: _lib.check(_lib.lib().bffc_kf_from_filter(plan.handle, _ptr(k), L, _ptr(kf_out), H, 0, _stream())), n=200))
if N <= 8192:
    dkf = torch.randn(H, 8192, 2, device=dev); dk = torch.empty(H, L, device=dev)
    print('bffc_dk_from_dkf            %.1f us' % timeit(lambda i: _lib.check(_lib.lib().bffc_dk_from_dkf(plan.handle, _ptr(dkf), _ptr(dk), L, H, _stream())), n=200))
print('module forward              %.1f us' % timeit(lambda i: mod(us[i % 3], k)))
ug = us[0].clone().requires_grad_(True); kg = k.clone().requires_grad_(True); dout = torch.randn_like(us[0])
def fb(i):
    ug.grad = None; kg.grad = None
    mod(ug, kg).backward(dout)
print('module forward + backward   %.1f us' % timeit(fb))
torch.cuda.synchronize(); t = time.perf_counter()
for i in range(200): mod(us[i % 3], k)
t_host = (time.perf_counter() - t) / 200 * 1e6
torch.cuda.synchronize()
print('module forward, host-side issue time %.1f us per call (asynchronous)' % t_host)

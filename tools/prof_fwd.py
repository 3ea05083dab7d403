"""Minimal launch sequence for ncu: N=8192 B=16 H=768 forward, kernel only (k_f pre-packed)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
from flashfftconv import FlashFFTConv, _lib
from flashfftconv.conv import _pack_kf, _ptr, _stream
N = int(os.environ.get('N', 8192)); B = int(os.environ.get('B', 16)); H = int(os.environ.get('H', 768)); L = int(os.environ.get('L', N))
iters = int(os.environ.get('ITERS', 3))
dev = torch.device('cuda')
mod = FlashFFTConv(N, dtype=torch.bfloat16); plan = mod.plan(dev)
u = torch.randn(B, H, L, device=dev).to(torch.bfloat16); k = torch.randn(H, L, device=dev) / L ** 0.5
kf = _pack_kf(mod, plan, k, 0); y = torch.empty_like(u)
for _ in range(iters):
    _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(u), _ptr(kf), None, None, _ptr(y), B, H, L, None, 0, _stream()))
torch.cuda.synchronize()
print('ok')

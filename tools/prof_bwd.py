"""Minimal launch sequence for ncu: backward through the C ABI (du + dk_f), k_f pre-packed.  Env as prof_fwd.py."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
from flashfftconv import FlashFFTConv, _lib
from flashfftconv.conv import _pack_kf, _ptr, _stream
N = int(os.environ.get('N', 8192)); B = int(os.environ.get('B', 16)); H = int(os.environ.get('H', 768)); L = int(os.environ.get('L', N))
iters = int(os.environ.get('ITERS', 3))
dev = torch.device('cuda')
mod = FlashFFTConv(N, dtype=torch.bfloat16); plan = mod.plan(dev)
u = torch.randn(B, H, L, device=dev).to(torch.bfloat16); dout = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
k = torch.randn(H, L, device=dev) / L ** 0.5
kfc = _pack_kf(mod, plan, k, 1); du = torch.empty_like(u)
dkf = torch.empty(H, mod.fft_size(dev), 2, device=dev)
nws = _lib.lib().bffc_workspace_bytes(plan.handle, B, H, L)
ws = torch.empty(nws, dtype=torch.uint8, device=dev) if nws else None
for _ in range(iters):
    _lib.check(_lib.lib().bffc_bwd(plan.handle, _ptr(dout), _ptr(u), None, _ptr(kfc), None, None, _ptr(du), _ptr(dkf), None, None,
                                   B, H, L, _ptr(ws), nws, _stream()))
torch.cuda.synchronize()
print('ok')

"""Minimal launch sequence for ncu: forward through the C ABI with k_f pre-packed.
Env: N (8192), B (16), H (768), L (N), GATED (0), ITERS (3)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
from flashfftconv import FlashFFTConv, _lib
from flashfftconv.conv import _pack_kf, _ptr, _stream
N = int(os.environ.get('N', 8192)); B = int(os.environ.get('B', 16)); H = int(os.environ.get('H', 768)); L = int(os.environ.get('L', N))
iters = int(os.environ.get('ITERS', 3)); gated = os.environ.get('GATED', '0') == '1'
dev = torch.device('cuda')
mod = FlashFFTConv(N, dtype=torch.bfloat16); plan = mod.plan(dev)
u = torch.randn(B, H, L, device=dev).to(torch.bfloat16); k = torch.randn(H, L, device=dev) / L ** 0.5
g = [torch.randn(B, H, L, device=dev).to(torch.bfloat16) for _ in range(2)] if gated else [None, None]
kf = _pack_kf(mod, plan, k, 0); y = torch.empty_like(u)
nws = _lib.lib().bffc_workspace_bytes(plan.handle, B, H, L)
ws = torch.empty(nws, dtype=torch.uint8, device=dev) if nws else None
for _ in range(iters):
    _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(u), _ptr(kf), _ptr(g[0]), _ptr(g[1]), _ptr(y), B, H, L, _ptr(ws), nws, _stream()))
torch.cuda.synchronize()
print('ok')

"""One module-level forward + backward per BASELINE config for ncu launch lists.  Env: W (c2|c3|c4|c5), ITERS (2)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
from flashfftconv import FlashFFTConv
import bench
N, B, H, L, gated = bench.shard_shape(os.environ.get('W', 'c3'), 1)
dev = torch.device('cuda')
conv = FlashFFTConv(N, dtype=torch.bfloat16).to(dev)
u = torch.randn(B, H, L, device=dev).to(torch.bfloat16).requires_grad_(True)
k = (torch.randn(H, L, device=dev) / L ** 0.5).requires_grad_(True)
g = [torch.randn(B, H, L, device=dev).to(torch.bfloat16).requires_grad_(True) for _ in range(2)] if gated else []
dout = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push('steps')
for _ in range(int(os.environ.get('ITERS', 2))):
    for t in [u, k] + g:
        t.grad = None
    conv(u, k, *g).backward(dout)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print('ok')

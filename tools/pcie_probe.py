"""PCIe probe: pinned host <-> device bandwidth, one direction at a time and both at once (two streams)."""
import torch
n = 201326592
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device='cuda'); d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d(): d_in.copy_(h_in, non_blocking=True)
def d2h(): h_out.copy_(d_out, non_blocking=True)
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
for name, fn in (('h2d', h2d), ('d2h', d2h), ('both', both)):
    ms = t(fn)
    print(f'{name}: {ms:.2f} ms for {n/1e6:.0f} MB per direction -> {n/ms/1e6:.1f} GB/s per direction')

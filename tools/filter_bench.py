"""Times the filter-side transforms (k -> k_f, dk_f -> dk) per BASELINE config, with the workspace (channel-group) size
as a parameter; `--once` runs one call per config for a profiler."""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'flash-fft-conv_b200')]
import __graft_entry__ as ge
ge.build()
import flashfftconv
from flashfftconv import _lib
from flashfftconv.conv import _ptr, _stream

CASES = {'c2': (8192, 768, 8192), 'c3': (32768, 1024, 16384), 'c4': (1048576, 128, 1048576), 'c5': (4194304, 8, 4194304),
         '64k': (65536, 256, 65536), '512k': (524288, 64, 524288)}
once = '--once' in sys.argv
names = [a for a in sys.argv[1:] if not a.startswith('--')] or list(CASES)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for name in names:
    N, H, Lk = CASES[name]
    mod = flashfftconv.FlashFFTConv(N, dtype=torch.bfloat16).cuda()
    plan = mod.plan(torch.device('cuda', 0))
    lib = _lib.lib()
    k = torch.randn(H, Lk, device='cuda')
    kf = torch.empty(H, N, dtype=torch.int32, device='cuda')
    dkf = torch.randn(H, N, 2, device='cuda')
    dk = torch.empty(H, Lk, device='cuda')
    rec = lib.bffc_filter_workspace_bytes(plan.handle, H)
    R = max(N // 8192, 1)
    per = 2 * (R // 2 + 1) * 65536
    full = (H + 1) // 2 * per
    sizes = [rec] if once or N == 8192 else sorted({per, rec // 2 // per * per or per, rec, min(full, 4 * rec), full})
    for nbytes in sizes:
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device='cuda')
        f = lambda: _lib.check(lib.bffc_kf_from_filter(plan.handle, _ptr(k), Lk, _ptr(kf), H, 0, _ptr(ws), nbytes, _stream()))
        g = lambda: _lib.check(lib.bffc_dk_from_dkf(plan.handle, _ptr(dkf), _ptr(dk), Lk, H, _ptr(ws), nbytes, _stream()))
        if once:
            f(); g(); torch.cuda.synchronize()
            continue
        tf, tg = timeit(f), timeit(g)
        io = (H * Lk * 4 + H * N * 4) / 1e9
        print(f'{name} N={N} H={H} ws={nbytes / 2**20:.0f} MB groups={-(-full // max(nbytes, 1)) if nbytes else 1}: '
              f'kf_from_filter {tf:.1f} us ({io / tf * 1e6:.0f} GB/s of k + k_f), dk_from_dkf {tg:.1f} us', flush=True)

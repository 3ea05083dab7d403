"""Same-box anchor: the UNMODIFIED reference kernels (baseline/_ref, built by baseline/build_ref.py for sm_100) next to
ours on identical inputs — output agreement (y, du, dk) and timings at the BASELINE configs.  Runs on the GPU box only
(`gpurun -- python baseline/run_ref.py`); writes gpurun_out/ref_gpu.json.  Nothing here is on the product path."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')


def load_reference():
    """The reference package as module `ref_flashfftconv` (its own name `flashfftconv` is ours on sys.path)."""
    import importlib.util
    if not os.path.isdir(os.path.join(REF, 'flashfftconv')):
        return None
    sys.path.insert(0, REF)                     # monarch_cuda*.so lives here
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location('ref_flashfftconv', os.path.join(REF, 'flashfftconv', '__init__.py'),
                                                  submodule_search_locations=[os.path.join(REF, 'flashfftconv')])
    mod = importlib.util.module_from_spec(spec)
    sys.modules['ref_flashfftconv'] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    import torch
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
    import __graft_entry__ as ge
    ge.build()
    from flashfftconv import FlashFFTConv
    ref = load_reference()
    if ref is None:
        print(json.dumps({'unavailable': 'baseline/_ref not built'}))
        return
    dev = torch.device('cuda')
    out = {'agreement': [], 'timing': []}

    def rel(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm()).item()

    def ev(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    # ---- agreement on identical inputs (reference test distribution, tests/test_flashfftconv.py:54-64; H % 16 == 0 for
    # ---- seqlen > 32768 as the reference requires, README.md:269)
    for N, B, H, L, gated in [(8192, 4, 32, 8192, False), (8192, 4, 32, 4096, True), (32768, 2, 16, 16384, True),
                              (32768, 2, 16, 32768, False), (1048576, 2, 16, 1048576, False)]:
        torch.manual_seed(0)
        u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
        k = torch.randn(H, L, device=dev) / L ** 0.5
        dout = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
        gates = [torch.randn(B, H, L, device=dev).to(torch.bfloat16) for _ in range(2)] if gated else []
        res = {}
        for name, cls in (('ours', FlashFFTConv), ('ref', ref.FlashFFTConv)):
            conv = cls(N, dtype=torch.bfloat16).to(dev)
            leaves = [t.clone().requires_grad_(True) for t in [u, k] + gates]
            y = conv(*leaves)
            y.backward(dout)
            torch.cuda.synchronize()
            res[name] = [y.detach()] + [t.grad for t in leaves]
        # fp32 truth on the GPU (torch.fft) for scale
        uf = (u.float() * gates[0].float()) if gated else u.float()
        yt = torch.fft.irfft(torch.fft.rfft(uf, n=N) * torch.fft.rfft(k, n=N), n=N)[..., :L]
        if gated:
            yt = yt * gates[1].float()
        names = ['y', 'du', 'dk'] + (['dpregate', 'dpostgate'] if gated else [])
        row = {'N': N, 'B': B, 'H': H, 'L': L, 'gated': gated,
               'ours_vs_ref': {n: rel(a, b) for n, a, b in zip(names, res['ours'], res['ref'])},
               'ours_vs_fp32': rel(res['ours'][0], yt), 'ref_vs_fp32': rel(res['ref'][0], yt)}
        out['agreement'].append(row)
        print(row, flush=True)

    # ---- timing at the BASELINE configs (forward, k_f included as both modules do it per call; fwd+bwd)
    for name, (N, B, H, L, gated) in {'c2': (8192, 16, 768, 8192, False), 'c3': (32768, 8, 1024, 16384, True),
                                      'c4': (1048576, 2, 128, 1048576, False), 'c5': (4194304, 8, 16, 4194304, False),
                                      'r256': (256, 64, 768, 256, True), 'r1k': (1024, 64, 768, 1024, True),
                                      'r4k': (4096, 64, 768, 4096, True), 'r8k': (8192, 64, 768, 8192, True)}.items():
        torch.manual_seed(1)
        u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
        k = torch.randn(H, L, device=dev) / L ** 0.5
        gates = [torch.randn(B, H, L, device=dev).to(torch.bfloat16) for _ in range(2)] if gated else []
        dout = torch.randn_like(u)
        row = {'config': name, 'N': N, 'B': B, 'H': H, 'L': L, 'gated': gated}
        for who, cls in (('ours', FlashFFTConv), ('ref', ref.FlashFFTConv)):
            try:
                conv = cls(N, dtype=torch.bfloat16).to(dev)
                n = 10 if N <= 32768 else 4
                row[who + '_fwd_ms'] = ev(lambda: conv(u, k, *gates), n)
                leaves = [t.clone().requires_grad_(True) for t in [u, k] + gates]

                def fb():
                    for t in leaves:
                        t.grad = None
                    conv(*leaves).backward(dout)
                row[who + '_fwd_bwd_ms'] = ev(fb, max(2, n // 2))
                del conv, leaves
            except Exception as e:
                row[who + '_error'] = f'{type(e).__name__}: {e}'[:200]
            torch.cuda.empty_cache()
        out['timing'].append(row)
        print(row, flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'ref_gpu.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()

"""GPU bring-up of the flop-lean inner kernel (fwd4_r16.cuh): compare the kernel's TMEM stage dumps with the dataflow
model (tests/kernel_model_r16.py), check full forwards against torch.fft, time it next to fwd3.
Run under gpurun with BFFC_INNER=4 (bring-up switch)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from flashfftconv import FlashFFTConv, _lib
from flashfftconv.conv import _pack_kf, _ptr, _stream
import kernel_model_r16 as km


def ref_fft_conv(u, k, n):
    l = u.size(-1)
    return torch.fft.irfft(torch.fft.rfft(u.float(), n=n) * torch.fft.rfft(k.float(), n=n), n=n)[..., :l]


def main():
    torch.manual_seed(0)
    N = 8192
    dev = torch.device('cuda')
    print(torch.cuda.get_device_name(0), 'BFFC_INNER =', os.environ.get('BFFC_INNER'), flush=True)
    mod = FlashFFTConv(N, dtype=torch.bfloat16)
    plan = mod.plan(dev)
    B, H, L = 2, 1, N
    u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
    k = torch.randn(H, L, device=dev) / (L ** 0.5)
    kf = _pack_kf(mod, plan, k, 0)
    y = torch.zeros_like(u)
    dump = torch.zeros(6, 128, 128, device=dev, dtype=torch.float32)
    rc = _lib.lib().bffc_debug_fwd_stages(plan.handle, _ptr(u), _ptr(kf), _ptr(y), B, H, L, _ptr(dump), 6, _stream())
    torch.cuda.synchronize()
    print('debug rc', rc, _lib.lib().bffc_last_error(), flush=True)
    kf_nat = torch.fft.fft(k.float(), n=N)[0].cpu().numpy().astype(np.complex128)
    x0 = u[0, 0].float().cpu().numpy().astype(np.float64); x1 = u[1, 0].float().cpu().numpy().astype(np.float64)
    y0m, y1m, st = km.model_fwd4(x0, x1, kf_nat, quant=True)
    d = dump.cpu().numpy().astype(np.float64)
    names = ['D1', 'D2', 'D3 spectrum', "D3'", "D2'", "D1' out"]
    for s in range(6):
        ref = st[s]; got = d[s]
        err = np.abs(got - ref).max(); sc = np.abs(ref).max()
        rl2 = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(f'stage {s} {names[s]:12s} max|ref|={sc:.4e} max err={err:.4e} rel-L2={rl2:.3e}', flush=True)
        if rl2 > 2e-2:
            # error map: rel-L2 per (32-lane block, 16-column block)
            em = [[np.linalg.norm(got[32 * a:32 * a + 32, 16 * b:16 * b + 16] - ref[32 * a:32 * a + 32, 16 * b:16 * b + 16]) /
                   (np.linalg.norm(ref[32 * a:32 * a + 32, 16 * b:16 * b + 16]) + 1e-30) for b in range(8)] for a in range(4)]
            for a in range(4):
                print('   lanes %3d..: ' % (32 * a) + ' '.join('%.2f' % e for e in em[a]))
            bad = np.argwhere(np.abs(got - ref) > 5e-2 * sc)
            print('   first bad (lane,col):', bad[:8].tolist(), ' n_bad', len(bad))
            for r in (0, 1, 9, 64):
                print(f'   got[{r},:6]', np.round(got[r, :6], 4), ' ref', np.round(ref[r, :6], 4))
            # is it a permutation of lanes / columns?  correlate column 0 of ref with every got column, lane 0 with every lane
            c0 = ref[:, 0]; cc = [abs(np.dot(c0, got[:, j])) / (np.linalg.norm(c0) * np.linalg.norm(got[:, j]) + 1e-30) for j in range(128)]
            print('   ref col 0 best matches got col', int(np.argmax(cc)), round(max(cc), 3))
            r0 = ref[0]; rr = [abs(np.dot(r0, got[i])) / (np.linalg.norm(r0) * np.linalg.norm(got[i]) + 1e-30) for i in range(128)]
            print('   ref lane 0 best matches got lane', int(np.argmax(rr)), round(max(rr), 3))
    yr = ref_fft_conv(u, k, N)
    print('unit fwd rel-L2 vs torch.fft', ((y.float() - yr).norm() / yr.norm()).item(),
          ' model says', np.linalg.norm(y0m - yr[0, 0].cpu().numpy()) / np.linalg.norm(yr[0, 0].cpu().numpy()), flush=True)
    ym = np.stack([y0m, y1m])
    print('unit fwd vs model rel-L2', np.linalg.norm(y[:, 0].float().cpu().numpy() - ym) / np.linalg.norm(ym), flush=True)
    for (B, H, L) in [(2, 4, N), (3, 5, N), (4, 16, N // 2), (16, 768, N)]:
        u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
        k = torch.randn(H, L, device=dev) / (L ** 0.5)
        y = mod(u, k)
        torch.cuda.synchronize()
        yr = ref_fft_conv(u, k, N)
        print(f'fwd B={B} H={H} L={L}: rel-L2 {((y.float() - yr).norm() / yr.norm()).item():.3e}', flush=True)
    # timing at C2 (kernel only, k_f pre-packed)
    B, H, L = 16, 768, N
    u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
    k = torch.randn(H, L, device=dev) / (L ** 0.5)
    kf = _pack_kf(mod, plan, k, 0); y = torch.empty_like(u)
    def kern():
        _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(u), _ptr(kf), None, None, _ptr(y), B, H, L, None, 0, _stream()))
    for _ in range(5):
        kern()
    torch.cuda.synchronize()
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            kern()
        e1.record(); torch.cuda.synchronize()
        print('C2 kernel %.1f us' % (e0.elapsed_time(e1) / 20 * 1000), flush=True)


if __name__ == '__main__':
    main()

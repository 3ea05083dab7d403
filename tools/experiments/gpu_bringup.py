"""GPU bring-up: compare the kernel's TMEM stage dumps with the float64/bf16 dataflow model, then check a
full forward against torch.fft and take a first timing.  Run under gpurun; writes gpurun_out/bringup.log."""
import ctypes, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from flashfftconv import FlashFFTConv, _lib
from flashfftconv.conv import _pack_kf, _ptr, _stream
import kernel_model_r128 as km

def ref_fft_conv(u, k, n):
    l = u.size(-1)
    u_f = torch.fft.fft(u.to(torch.float32), n=n)
    k_f = torch.fft.fft(k.to(torch.float32), n=n)
    return torch.fft.ifft(u_f * k_f, n=n).real.to(u.dtype)[..., :l]

def main():
    torch.manual_seed(0)
    N = 8192
    dev = torch.device('cuda')
    print(torch.cuda.get_device_name(0), flush=True)
    mod = FlashFFTConv(N, dtype=torch.bfloat16)
    plan = mod.plan(dev)
    # ---- stage dumps on one unit
    B, H, L = 2, 1, N
    u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
    k = torch.randn(H, L, device=dev) / (L ** 0.5)
    kf = _pack_kf(mod, plan, k, 0)
    y = torch.zeros_like(u)
    dump = torch.zeros(4, 128, 128, device=dev, dtype=torch.float32)
    rc = _lib.lib().bffc_debug_fwd_stages(plan.handle, _ptr(u), _ptr(kf), _ptr(y), B, H, L, _ptr(dump), 4, _stream())
    torch.cuda.synchronize()
    print('debug rc', rc, flush=True)
    kf_nat = torch.fft.fft(k.float(), n=N)[0].cpu().numpy().astype(np.complex128)
    x0 = u[0, 0].float().cpu().numpy().astype(np.float64); x1 = u[1, 0].float().cpu().numpy().astype(np.float64)
    y0m, y1m, st = km.model_fwd(x0, x1, kf_nat, quant=True)
    d = dump.cpu().numpy().astype(np.float64)
    names = ['D1 outer DFT', 'D2 spectrum', 'D3 inverse64', 'D4 out']
    for s in range(4):
        ref = st[s]; got = d[s]
        err = np.abs(got - ref).max(); sc = np.abs(ref).max()
        print(f'stage {s} {names[s]:14s} max|ref|={sc:.4e} max err={err:.4e} rel={err/sc:.3e}', flush=True)
        if err / sc > 5e-2:
            # diagnostics: which rows / cols are off
            bad = np.argwhere(np.abs(got - ref) > 5e-2 * sc)
            print('   first bad (lane,col):', bad[:8].tolist(), ' n_bad', len(bad))
            print('   got[0,:8]', got[0, :8], '\n   ref[0,:8]', ref[0, :8])
            print('   got[1,:8]', got[1, :8], '\n   ref[1,:8]', ref[1, :8])
    yr = ref_fft_conv(u, k, N).float()
    e = (y.float() - yr).norm() / yr.norm()
    print('unit fwd rel-L2 vs torch.fft', e.item(), flush=True)
    # ---- full forward parity, several shapes
    for (B, H, L) in [(2, 4, N), (3, 5, N), (4, 16, N // 2), (16, 768, N)]:
        u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
        k = torch.randn(H, L, device=dev) / (L ** 0.5)
        y = mod(u, k)
        torch.cuda.synchronize()
        yr = ref_fft_conv(u, k, N).float()
        e = ((y.float() - yr).norm() / yr.norm()).item()
        m = ((y.float() - yr).abs().max() / yr.abs().max()).item()
        print(f'fwd B={B} H={H} L={L}: rel-L2 {e:.3e} max-rel {m:.3e}', flush=True)
    # ---- timing at config 2 (kernel only, k_f precomputed)
    B, H, L = 16, 768, N
    u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
    k = torch.randn(H, L, device=dev) / (L ** 0.5)
    kf = _pack_kf(mod, plan, k, 0)
    y = torch.empty_like(u)
    def run():
        _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(u), _ptr(kf), None, None, _ptr(y), B, H, L, None, 0, _stream()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10): run()
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print(f'C2 kernel-only: {ms*1e3:.1f} us/step  {B*H/ms*1e3:.3e} convs/s  {(4*L*B*H + 4*N*H)/ms/1e6:.1f} GB/s', flush=True)

if __name__ == '__main__':
    main()

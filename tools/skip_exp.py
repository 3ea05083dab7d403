"""Experiment: fused forward kernel (C2 shape) with parts disabled (env BFFC_SKIP: 1 = no MMAs, 2 = no pass arithmetic,
4 = no TMA stores).  Prints average kernel time and the SM clock / power sampled while the loop runs."""
import os, sys, subprocess, threading, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
from flashfftconv import FlashFFTConv, _lib
from flashfftconv.conv import _pack_kf, _ptr, _stream
N = 8192; B = int(os.environ.get('B', 16)); H = int(os.environ.get('H', 768)); L = N
dev = torch.device('cuda')
mod = FlashFFTConv(N, dtype=torch.bfloat16).cuda(); plan = mod.plan(dev)
us = [torch.randn(B, H, L, device=dev).to(torch.bfloat16) for _ in range(3)]
k = torch.randn(H, L, device=dev) / L ** 0.5
kf = _pack_kf(mod, plan, k, 0); y = torch.empty_like(us[0])
def kern(i): _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(us[i % 3]), _ptr(kf), None, None, _ptr(y), B, H, L, None, 0, _stream()))
for i in range(20): kern(i)
torch.cuda.synchronize()
samples = []
def sampler():
    p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-lms', '50'],
                         stdout=subprocess.PIPE, text=True)
    t0 = time.time()
    for line in p.stdout:
        samples.append(line.strip())
        if time.time() - t0 > 1.6: break
    p.terminate()
th = threading.Thread(target=sampler); th.start()
time.sleep(0.5)
n = int(os.environ.get('ITERS', 6000))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n): kern(i)
e1.record(); torch.cuda.synchronize()
th.join()
mid = samples[len(samples) // 2:]
print('SKIP=%s  kernel %.1f us   clocks/power (late samples): %s' % (os.environ.get('BFFC_SKIP', '0'), e0.elapsed_time(e1) / n * 1000, ' | '.join(mid[-6:])))

"""Sustained (power-capped steady state) vs burst time of the C2 forward kernel: LOOPS back-to-back launches, SM clock and
power sampled with nvidia-smi meanwhile.  BFFC_INNER=4 selects the flop-lean inner kernel (bring-up switch)."""
import os, subprocess, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))
from flashfftconv import FlashFFTConv, _lib
from flashfftconv.conv import _pack_kf, _ptr, _stream
N, B, H = 8192, 16, 768
dev = torch.device('cuda')
mod = FlashFFTConv(N, dtype=torch.bfloat16); plan = mod.plan(dev)
u = torch.randn(B, H, N, device=dev).to(torch.bfloat16); k = torch.randn(H, N, device=dev) / N ** 0.5
kf = _pack_kf(mod, plan, k, 0); y = torch.empty_like(u)
def kern():
    _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(u), _ptr(kf), None, None, _ptr(y), B, H, N, None, 0, _stream()))
samples = []
def smi():
    p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        samples.append((time.time(), line.strip()))
        if stop[0]:
            p.terminate(); break
stop = [False]
t = threading.Thread(target=smi, daemon=True); t.start()
for _ in range(5): kern()
torch.cuda.synchronize(); time.sleep(1.0)
def timed(n):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): kern()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
print('inner =', os.environ.get('BFFC_INNER', '3 (fwd3)'))
print('burst  20 launches: %.1f us' % timed(20))
time.sleep(1.0)
t0 = time.time()
for rep in range(4):
    print('loop 4000 launches: %.1f us' % timed(4000), flush=True)
t1 = time.time()
stop[0] = True
inside = [s for (ts, s) in samples if t0 + 0.5 <= ts <= t1]
clk = sorted(float(s.split(',')[0]) for s in inside); pw = [float(s.split(',')[1]) for s in inside]
if clk: print('under load: sm clock median %.0f MHz (min %.0f), power mean %.0f W max %.0f W, %d samples' % (clk[len(clk)//2], clk[0], sum(pw)/len(pw), max(pw), len(clk)))

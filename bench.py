#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the fused FFT-convolution hot path.

  python bench.py --gpus N --steps K --warmup W            # our sm_100a engine
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's CPU path
                                                           # (tests/test_flashfftconv.py:5-13 oracle port)

One "step" = one pass of the hot path over one synthetic batch of BASELINE.json's configs[1]
(N=8192, B=16, H=768, bf16, ungated, L=N) per GPU.  Prints ONE JSON line (rank 0).  Multi-GPU runs shard
B x H with no data-path collective (every (b,h) convolution is independent): each rank owns its own H=768
channel block (weak scaling); NCCL is used only for the barrier / max-over-ranks timing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'flash-fft-conv_b200'))

import torch  # noqa: E402

_OUT_FD = 1

WORKLOADS = {
    # name: (N, B, H, L, gated)
    'c2': (8192, 16, 768, 8192, False),       # BASELINE.json configs[1]: M2-BERT dims, the metric's config
    'c3': (32768, 8, 1024, 16384, True),      # configs[2]: Hyena-style, gated, implicit 2x causal padding
    'c4': (1048576, 2, 128, 1048576, False),  # configs[3]: HyenaDNA long range (per GPU here; see config.sharding)
    'c5': (4194304, 8, 8, 4194304, False),    # configs[4] per-GPU shard: B=8, H=64/8
}


def algorithmic_bytes(N, B, H, L, gated):
    """fwd ungated 4L per conv (+4L gates when gated) plus k_f once per channel (4N)."""
    return (8 if gated else 4) * L * B * H + 4 * N * H


def issued_tensor_flops(N, B, H):
    """Matmul flops the inner 8192-point kernel issues: 25.2 MFLOP per unit (a pair of sequences of one
    channel; an odd batch still runs a full unit), N/8192 units per pair for the composite sizes; the outer radix-128
    stage of the 1M+ sizes adds 2 x 128 x 256 x 2 flops per complex column.  Sizes below 8192 share one unit between
    4096/N batch pairs."""
    unit = 2.0 * 128 * 128 * (2 * 256 + 2 * 128)
    ne = max(N, 8192)
    seg = 4096 // N if N < 8192 else 1
    groups = (B + 2 * seg - 1) // (2 * seg)
    flops = groups * H * (ne // 8192) * unit
    if N >= (1 << 20):
        flops += ((B + 1) // 2) * H * 2.0 * (2.0 * 128 * 128 * 256) * (N // 128 // 64)
    return flops


def reference_tensor_flops(N, B, H):
    """SURVEY.md §8(d) 'algorithmic flops per conv' of the reference's own factorisation, forward."""
    per_conv = {8192: 6.29e6, 32768: 41.9e6, 1 << 20: 1.88e9, 1 << 22: 10.7e9}.get(N)
    return per_conv * B * H if per_conv else None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the GPU is under the benchmark load
    (B200_PROFILING.md 'clocks line'): one persistent `nvidia-smi -lms 100`, samples are time-stamped and only
    those taken inside a marked load window are summarised."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []          # (t, fields)
        self.windows = []          # (t0, t1) under load
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.strip().split(',')]
                if len(f) >= 7:
                    self.samples.append((time.time(), f))
        except Exception:
            pass

    def stop(self):
        try:
            if self.proc:
                self.proc.terminate()
        except Exception:
            pass

    def summary(self):
        inside = [f for (t, f) in self.samples if any(a <= t <= b for a, b in self.windows)]
        if not inside:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no nvidia-smi sample under load']}
        sm = sorted(float(s[0]) for s in inside)
        reasons = []
        for i, name in [(3, 'hw_slowdown'), (4, 'hw_thermal_slowdown'), (5, 'sw_thermal_slowdown'), (6, 'sw_power_cap')]:
            if any(s[i].lower().startswith('active') for s in inside):
                reasons.append(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(inside[0][1]), 'reasons': reasons,
                'samples_under_load': len(inside),
                'power_w_max': max(float(s[2]) for s in inside if s[2].replace('.', '', 1).isdigit())}


def sample_shape(N):
    n = max(2, (1 << 21) // N)
    return (4, n // 4) if n >= 8 else (2, n // 2)


def cpu_baseline(N, L, gated, budget_s=12.0):
    """The reference's CPU path (oracle port of tests/test_flashfftconv.py:5-13) on the host cores,
    on a bounded sample of the same workload: S convolutions of the true N / L, all threads."""
    from oracle.fftconv_oracle import ref_fft_conv, ref_fft_conv_gated
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs, Hs = sample_shape(N)                          # bounded sample: ~2M points per call
    g = torch.Generator().manual_seed(0)
    u = torch.randn(Bs, Hs, L, generator=g).to(torch.bfloat16)
    k = torch.randn(Hs, L, generator=g) / L ** 0.5
    if gated:
        pg = torch.randn(Bs, Hs, L, generator=g).to(torch.bfloat16)
        qg = torch.randn(Bs, Hs, L, generator=g).to(torch.bfloat16)
        fn = lambda: ref_fft_conv_gated(u, k, pg, qg, N)
    else:
        fn = lambda: ref_fft_conv(u, k, N)
    for _ in range(5):          # thread pool / FFT plan warm-up (first calls are 10x slower)
        fn()
    t0 = time.perf_counter(); n = 0
    while True:
        fn(); n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 200:
            break
    return {'value': Bs * Hs * n / dt, 'unit': 'convs/s', 'cores': cores, 'kind': 'port',
            'sample': f'{n} calls x {Bs * Hs} convs (B={Bs},H={Hs}) at N={N}, L={L}, fp32 torch.fft on CPU'}


def dist_setup(n_gpus):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return rank, world, local


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    N, B, H, L, gated = WORKLOADS[args.workload]
    from oracle.fftconv_oracle import ref_fft_conv
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs, Hs = sample_shape(N)
    g = torch.Generator().manual_seed(0)
    u = torch.randn(Bs, Hs, L, generator=g).to(torch.bfloat16)
    k = torch.randn(Hs, L, generator=g) / L ** 0.5
    for _ in range(args.warmup):
        ref_fft_conv(u, k, N)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_fft_conv(u, k, N)
    dt = time.perf_counter() - t0
    val = Bs * Hs * args.steps / dt
    sample = f'each step = {Bs * Hs} convs (B={Bs},H={Hs}) of the N={N}, L={L} workload; fp32 torch.fft, {cores} threads'
    emit(json.dumps({
        'impl': 'reference', 'metric': 'fftconv_fwd_convs_per_sec', 'value': val, 'unit': 'convs/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: N={N} B={B} H={H} L={L} ungated (bounded CPU sample)'},
        'cpu_baseline': {'value': val, 'unit': 'convs/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'convs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}), _OUT_FD)


def run_ours(args):
    rank, world, local = dist_setup(args.gpus)
    import __graft_entry__ as ge
    ge.build()
    from flashfftconv import FlashFFTConv, _lib
    from flashfftconv.conv import _pack_kf, _ptr, _stream
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    N, B, H, L, gated = WORKLOADS[args.workload]
    torch.manual_seed(1234 + rank)
    conv = FlashFFTConv(N, dtype=torch.bfloat16).to(dev)
    plan = conv.plan(dev)
    u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
    k = torch.randn(H, L, device=dev) / L ** 0.5
    gates = [torch.randn(B, H, L, device=dev).to(torch.bfloat16) for _ in range(2)] if gated else []
    convs_per_step = B * H

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- (1) device-resident whole step through the public forward (k -> k_f -> fused conv)
    for _ in range(args.warmup):
        y = conv(u, k, *gates)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    t_load0 = time.time()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        y = conv(u, k, *gates)
        launches += 1 + _lib.lib().bffc_last_launch_count()          # kf_pack + conv kernels (our kernels only)
    e1.record()
    barrier()
    step_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)

    # ---- (2) dominant kernel alone (k_f pre-packed), CUDA events on the launching stream -> roofline
    kf = _pack_kf(conv, plan, k, 0)
    yk = torch.empty_like(u)
    ws_bytes = _lib.lib().bffc_workspace_bytes(plan.handle, B, H, L)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    g0 = _ptr(gates[0]) if gated else None
    g1 = _ptr(gates[1]) if gated else None

    def kern():
        _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(u), _ptr(kf), g0, g1, _ptr(yk), B, H, L, _ptr(ws), ws_bytes,
                                       _stream()))
    for _ in range(max(3, args.warmup)):
        kern()
    torch.cuda.synchronize()
    k0 = torch.cuda.Event(enable_timing=True); k1 = torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        kern()
    k1.record(); torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / args.steps
    if sampler:
        # the timed regions last milliseconds; keep the identical kernel running ~1.5 s more so that
        # nvidia-smi (100 ms period) sees clocks and throttle reasons under this load
        t_end = time.time() + 1.5
        while time.time() < t_end:
            for _ in range(50):
                kern()
            torch.cuda.synchronize()
        sampler.windows.append((t_load0, time.time()))
        time.sleep(0.15)
        sampler.stop()

    # ---- (2b) forward + backward through autograd (du, dk), device resident
    ug = u.clone().requires_grad_(True); kg = k.clone().requires_grad_(True)
    gg = [g.clone().requires_grad_(True) for g in gates]
    dout = torch.randn_like(u)
    def fb():
        ug.grad = None; kg.grad = None
        for g in gg:
            g.grad = None
        conv(ug, kg, *gg).backward(dout)
    for _ in range(3):
        fb()
    barrier()
    f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        fb()
    f1.record()
    barrier()
    fb_ms = max_over_ranks(f0.elapsed_time(f1) / args.steps)

    # ---- (3) end to end through the public API with HOST buffers (pinned), copies inside the timed region
    u_h = u.cpu().pin_memory(); k_h = k.cpu().pin_memory()
    g_h = [g.cpu().pin_memory() for g in gates]
    y_h = torch.empty_like(u_h).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        # public host-buffer call: k -> device, k_f, then u (and gates) host -> device, conv, y device -> host,
        # pipelined over batch chunks inside bffc_fwd_host (include/bffc.h)
        conv.forward_host(u_h, k_h, *g_h, out=y_h, device=dev)
    e2e_step()
    barrier()
    s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(e2e_steps):
        e2e_step()
    s1.record()
    barrier()
    e2e_ms = max_over_ranks(s0.elapsed_time(s1) / e2e_steps)

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    bf16_peak = float(peaks.get('bf16_tflops', peaks.get('bf16_tfs', 0)) or 0)
    if not bf16_peak:
        bf16_peak = next((float(v) for k_, v in peaks.items() if 'bf16' in k_.lower() and isinstance(v, (int, float))), 1640.0)
    tflops_issued = issued_tensor_flops(N, B, H)
    ref_flops = reference_tensor_flops(N, B, H)
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s'
    abytes = algorithmic_bytes(N, B, H, L, gated)
    achieved = abytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json'))).get(args.workload)
    except Exception:
        pass
    out = {
        'metric': 'fftconv_fwd_convs_per_sec', 'value': convs_per_step * world / (step_ms * 1e-3), 'unit': 'convs/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': step_ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: FlashFFTConv({N}, bf16) fwd, B={B} H={H} L={L} '
                               f'{"gated" if gated else "ungated"} per GPU (BASELINE.json configs); '
                               f'step = k->k_f (bffc_kf_from_filter for seqlen <= 8192, else torch.fft.rfft + bffc_kf_pack_rfft) + conv kernels',
                   'l2': f'inputs+outputs {abytes / 1e6:.0f} MB per step exceed the 126 MB L2 (no flush needed)',
                   'sharding': 'B x H sharded over ranks, no data-path collective'},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
                     'traffic': traffic,
                     'kernel': 'bffc::r128::fwd3_kernel' if N == 8192 and not gated else 'bffc_fwd native path (outer stages / fold + bffc::r128 fused kernel)',
                     'kernel_ms': kern_ms,
                     'algorithmic_bytes': abytes, 'peak_source': peak_src,
                     'kernel_convs_per_sec': convs_per_step / (kern_ms * 1e-3)},
        # the tighter roofline at 8192 under this factorisation (report both): flops actually issued
        # to the tensor pipe per launch vs the measured dense bf16 peak (itself power limited on this board)
        'roofline_tensor': {'bound': 'tensor', 'achieved': tflops_issued / (kern_ms * 1e-3) / 1e12, 'peak': bf16_peak,
                            'unit': 'TFLOP/s', 'frac': tflops_issued / (kern_ms * 1e-3) / 1e12 / bf16_peak,
                            'issued_flops': tflops_issued, 'reference_factorisation_flops': ref_flops,
                            'note': 'radix 128 x 64 issues 2x the matmul flops of the reference split (DESIGN.md §6)'},
        'e2e': {'value': convs_per_step * world / (e2e_ms * 1e-3), 'unit': 'convs/s',
                'h2d_bytes_per_step': u_h.numel() * 2 * (3 if gated else 1) + k_h.numel() * 4,
                'd2h_bytes_per_step': y_h.numel() * 2,
                'ms_per_step': e2e_ms},
        'fwd_bwd': {'value': convs_per_step * world / (fb_ms * 1e-3), 'unit': 'convs/s', 'ms_per_step': fb_ms,
                    'note': 'autograd fwd+bwd (du, dk) through FlashFFTConv, inputs resident'},
        'gpu_launches': launches,
        'clocks': sampler.summary() if sampler else None,
    }
    out['cpu_baseline'] = cpu_baseline(N, L, gated) if world == 1 else None
    emit(json.dumps(out), _OUT_FD)


def main():
    # libraries (NCCL prints its version banner) must not pollute the ONE JSON line on stdout: send fd 1 to stderr
    # while the benchmark runs and restore it for the final print
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        _main(saved_fd)
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)


def emit(line, saved_fd):
    os.write(saved_fd, (line + '\n').encode())


def _main(saved_fd):
    global _OUT_FD
    _OUT_FD = saved_fd
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
    if int(os.environ.get('WORLD_SIZE', '1')) > 1 and args.impl == 'ours':
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

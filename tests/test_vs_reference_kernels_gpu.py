"""Agreement with the reference's OWN CUDA kernels on identical inputs (north_star: "outputs match the reference's own
kernels").  The reference extension is built by baseline/build_ref.py into the git-ignored baseline/_ref/ (it travels to
the GPU box with gpurun); the test is skipped when it is absent.  Bar: our result is at least as close to the fp32
truth as the reference's, and the two agree within the sum of their distances to it."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'flashfftconv')) and any(
    f.startswith('monarch_cuda') and f.endswith('.so') for f in os.listdir(os.path.join(ROOT, 'baseline', '_ref')))


@pytest.fixture(scope='module')
def both():
    import __graft_entry__ as ge
    ge.build()
    import flashfftconv
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import run_ref
    return flashfftconv, run_ref.load_reference()


@pytest.mark.skipif(not HAVE_REF, reason='baseline/_ref (reference CUDA extension) not built')
@pytest.mark.parametrize('N,B,H,L,gated', [(8192, 4, 32, 8192, False), (8192, 2, 16, 4096, True), (32768, 2, 16, 16384, True),
                                           (1048576, 2, 16, 1048576, False)])
def test_matches_reference_kernels(both, N, B, H, L, gated):
    ours, ref = both
    dev = torch.device('cuda')
    torch.manual_seed(0)
    u = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
    k = torch.randn(H, L, device=dev) / L ** 0.5
    dout = torch.randn(B, H, L, device=dev).to(torch.bfloat16)
    gates = [torch.randn(B, H, L, device=dev).to(torch.bfloat16) for _ in range(2)] if gated else []
    res = {}
    for name, cls in (('ours', ours.FlashFFTConv), ('ref', ref.FlashFFTConv)):
        conv = cls(N, dtype=torch.bfloat16).to(dev)
        leaves = [t.clone().requires_grad_(True) for t in [u, k] + gates]
        conv(*leaves).backward(dout)
        y = conv(*[t.detach() for t in leaves])
        res[name] = [y] + [t.grad for t in leaves]
    # fp32 truth through autograd on the GPU (torch.fft), same operator as the oracle
    lf = [t.float().clone().requires_grad_(True) for t in [u, k] + gates]
    x = lf[0] * lf[2] if gated else lf[0]
    yt = torch.fft.irfft(torch.fft.rfft(x, n=N) * torch.fft.rfft(lf[1], n=N), n=N)[..., :L]
    if gated:
        yt = yt * lf[3]
    yt.backward(dout.float())
    truth = [yt.detach()] + [t.grad for t in lf]
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
    for i, name in enumerate(['y', 'du', 'dk', 'dpregate', 'dpostgate'][: len(truth)]):
        e_ours, e_ref, e_x = rel(res['ours'][i], truth[i]), rel(res['ref'][i], truth[i]), rel(res['ours'][i], res['ref'][i])
        assert e_ours <= 1e-2, (name, e_ours)
        if e_ref <= 2e-2:       # the reference's own gated 32K du is off by O(1) on this box (profiles/r2_ref_gpu.md)
            assert e_x <= e_ours + e_ref + 1e-3, (name, e_x, e_ours, e_ref)

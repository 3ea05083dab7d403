"""CPU test of the multi-GPU host logic: world_size-2 gloo, channel sharding + harness gather.  The compute
function is the oracle here (no GPU in this container); on the GPU box the same code runs with FlashFFTConv."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, H, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, 'flash-fft-conv_b200'))
    from flashfftconv import parallel
    from oracle.fftconv_oracle import ref_fft_conv, make_inputs
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    N = 1024
    d = make_inputs(2, H, N, N, torch.float32, seed=3)
    y = parallel.sharded_conv(lambda u, k: ref_fft_conv(u, k, N), d['u'], d['k'])
    full = ref_fft_conv(d['u'], d['k'], N)
    ok = torch.equal(y, full)
    h0, h1 = parallel.channel_range(H, world, rank)
    q.put((rank, ok, h0, h1))
    dist.destroy_process_group()


@pytest.mark.parametrize('H', [16, 7])
def test_channel_sharding_gloo_world2(H):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == H     # contiguous cover of the channels


def test_channel_range_cover():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'flash-fft-conv_b200'))
    from flashfftconv.parallel import channel_range
    for H in (1, 7, 64, 768):
        for w in (1, 2, 4, 8):
            r = [channel_range(H, w, i) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == H
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1

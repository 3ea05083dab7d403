"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/bffc.h declares,
fails loudly without a GPU, and the Python mirror keeps the reference's constructor contract."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'bffc.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bffc_[a-z_0-9]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from flashfftconv import _lib
    return _lib


def test_header_symbols_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 10
    l = ctypes.CDLL(lib.LIB_PATH)
    for s in syms:
        assert hasattr(l, s), f'{s} declared in bffc.h but not exported'
    assert sorted(lib.SYMBOLS.keys()) == syms, 'ctypes binding out of sync with bffc.h'


def test_abi_version_and_supported(lib):
    l = lib.lib()
    assert l.bffc_abi_version() == 3
    assert l.bffc_supported(8192, lib.BFFC_DTYPE_BF16) == 1
    assert l.bffc_supported(8191, lib.BFFC_DTYPE_BF16) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_fails_loudly_without_gpu(lib):
    l = lib.lib()
    h = ctypes.c_void_p(0)
    rc = l.bffc_plan_create(ctypes.byref(h), 8192, lib.BFFC_DTYPE_BF16)
    assert rc == 3 and b'no CUDA device' in l.bffc_last_error()
    from flashfftconv import FlashFFTConv
    m = FlashFFTConv(8192, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 8192, dtype=torch.bfloat16), torch.zeros(1, 8192))


def test_module_contract(lib):
    from flashfftconv import FlashFFTConv
    with pytest.raises(AssertionError):
        FlashFFTConv(8192, dtype=torch.float32)                # conv.py:74
    with pytest.raises(NotImplementedError):
        FlashFFTConv(1000, dtype=torch.bfloat16)               # conv.py:550-551
    m = FlashFFTConv(8192, dtype=torch.bfloat16)
    assert isinstance(m, torch.nn.Module) and m.seqlen == 8192


def test_module_copies_pickles_and_loads_reference_checkpoints(lib):
    """The module must behave like the reference's under copy.deepcopy (EMA copies), pickle / torch.save(model) and
    load_state_dict(strict=True) of a checkpoint written by the reference, whose constant tables are persistent
    buffers (reference conv.py:89-92).  Native handles are per process and rebuilt lazily."""
    import copy
    import io
    import pickle
    from flashfftconv import FlashFFTConv

    class Stub:                      # stands in for a live native plan (ctypes pointers cannot be pickled)
        handle = ctypes.c_void_p(1234)
    m = FlashFFTConv(8192, dtype=torch.bfloat16)
    m._plans[('cuda', 0)] = Stub()
    m._host_ws[('cuda', 1)] = torch.zeros(4)
    m.eval()
    c = copy.deepcopy(m)
    assert c is not m and c.seqlen == 8192 and c.dtype == torch.bfloat16 and c._plans == {} and not c.training
    r = pickle.loads(pickle.dumps(m))
    assert r.seqlen == 8192 and r._plans == {} and r._host_ws == {} and r._kf_cache is None
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    assert torch.load(buf, weights_only=False).seqlen == 8192
    m._plans.clear()                 # the stub must not reach bffc_plan_destroy

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(2, 2)
            self.flashfftconv = FlashFFTConv(8192, dtype=torch.bfloat16)
    net = Net()
    sd = net.state_dict()
    assert not any(k.startswith('flashfftconv.') for k in sd)
    for name, shape in [('f_32_fft', (32, 32, 2)), ('f_16_ifft', (16, 16, 2)), ('twiddle_factors_fft_32_256', (32, 256, 2))]:
        sd['flashfftconv.' + name] = torch.zeros(shape, dtype=torch.bfloat16)   # what a reference checkpoint carries
    res = net.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'flash-fft-conv_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                assert 'oracle' not in open(os.path.join(d, f)).read().replace('the oracle', ''), f

"""Host-side mirror of the reference operator API for the fused FFT-convolution path.

Drop-in for `flashfftconv.FlashFFTConv` (reference flashfftconv/conv.py:71-560): same constructor
`FlashFFTConv(seqlen, dtype=torch.float16, use_32_butterfly=True)`, same
`forward(u, k, pregate=None, postgate=None)`, same autograd contract
(`backward -> (du, dk, None[, dpregate, dpostgate])`, conv.py:1822, :3939).

All arithmetic on the hot path happens in libbffc.so (hand-written sm_100a CUDA, C ABI in
include/bffc.h).  PyTorch is used for device memory and streams only: the filter-side transforms
(k -> k_f, reference conv.py:575 + :640; dk_f -> dk, conv.py:1817-1820) are the library's own fp32 FFT launches
for every supported size (bffc_kf_from_filter / bffc_dk_from_dkf) — no library FFT call is left in this module.

The filter spectrum in engine order is what forward keeps for backward (the reference keeps its permuted
k_f, conv.py:587-588), so a training step transforms the filter once; in eval mode it is additionally cached
across calls for as long as the SAME tensor object `k` is unmodified (identity + `_version`), which is the
inference situation of the reference's examples (one fixed filter, many inputs).
"""
import ctypes
import weakref

import torch

from . import _lib

_DT = {torch.bfloat16: _lib.BFFC_DTYPE_BF16, torch.float16: _lib.BFFC_DTYPE_FP16}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    """Raw handle of the current stream of the current device.  torch.cuda.current_stream() builds a Stream object
    (~13 us per call, measured: tools/host_overhead.py); the raw query underneath it is what Triton's launcher uses too."""
    try:
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    except AttributeError:                                         # private API moved: fall back to the public path
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Plan:
    """Owns one bffc_plan per (seqlen, dtype, device)."""

    def __init__(self, seqlen, dtype, device):
        self.handle = ctypes.c_void_p(0)
        self.device = device
        with torch.cuda.device(device):
            _lib.check(_lib.lib().bffc_plan_create(ctypes.byref(self.handle), int(seqlen), _DT[dtype]))
        # constants of the plan and memoised size queries: the per-call host path is a handful of ctypes calls, and at
        # C2 a forward step is 0.16 ms of GPU work — Python overhead shows up as launch gaps (8 ranks per box: 0.18 ms)
        self.fft_size = _lib.lib().bffc_fft_size(self.handle)
        self.length_multiple = _lib.lib().bffc_length_multiple(self.handle)
        self._ws_bytes = {}
        self._filter_ws_bytes = {}

    def workspace_bytes(self, B, H, L, gated, backward):
        key = (B, H, L, gated, backward)
        n = self._ws_bytes.get(key)
        if n is None:
            n = self._ws_bytes[key] = _lib.lib().bffc_workspace_bytes_ex(self.handle, B, H, L, int(gated), int(backward))
        return n

    def filter_workspace_bytes(self, H):
        n = self._filter_ws_bytes.get(H)
        if n is None:
            n = self._filter_ws_bytes[H] = _lib.lib().bffc_filter_workspace_bytes(self.handle, int(H))
        return n

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().bffc_plan_destroy(self.handle)
        except Exception:
            pass


class FlashFFTConv(torch.nn.Module):
    def __init__(self, seqlen, dtype=torch.float16, use_32_butterfly=True):
        super().__init__()
        assert dtype == torch.bfloat16 or dtype == torch.float16      # conv.py:74
        self.seqlen = int(seqlen)
        self.dtype = dtype
        self.use_32_butterfly = use_32_butterfly                     # accepted for API parity; no effect here
        if not _lib.lib().bffc_supported(self.seqlen, _DT[dtype]):
            raise NotImplementedError(f'seqlen {seqlen} not supported')   # conv.py:550-551
        self._reset_runtime_state()

    # ---- runtime state (native handles, device scratch, caches) is per process and rebuilt lazily: it must not
    # ---- travel through copy.deepcopy / pickle / torch.save(model) (ctypes pointers cannot be pickled, and two copies
    # ---- of a plan handle would be destroyed twice)
    def _reset_runtime_state(self):
        # plain-dict writes: torch.nn.Module.__setattr__ costs ~4 us per assignment, the hot path makes several per call
        d = self.__dict__
        d['_plans'] = {}
        d['_host_ws'] = {}
        d['_kf_cache'] = None          # (weakref(k), k._version, device, kf_engine)
        d['last_launches'] = 0         # kernels enqueued by the most recent forward / backward (bench.py)

    def __getstate__(self):
        state = self.__dict__.copy()
        for key in ('_plans', '_host_ws', '_kf_cache'):
            state.pop(key, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._reset_runtime_state()

    def __deepcopy__(self, memo):
        new = type(self)(self.seqlen, self.dtype, self.use_32_butterfly)
        new.train(self.training)
        memo[id(self)] = new
        return new

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # The reference module registers its DFT / twiddle tables as persistent buffers (conv.py:89-92 and every size
        # branch), so its checkpoints carry `<prefix>f_32_fft`, `<prefix>twiddle_factors_fft_16_16`, ...  This module
        # owns no tensors (the tables live in the native plan): accept and drop those entries so that
        # load_state_dict(strict=True) of a reference checkpoint succeeds.
        for key in [k for k in state_dict if k.startswith(prefix)]:
            del state_dict[key]
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def plan(self, device):
        key = (device.type, device.index)
        if key not in self._plans:
            self._plans[key] = _Plan(self.seqlen, self.dtype, device)
        return self._plans[key]

    def fft_size(self, device):
        """FFT size of the engine: seqlen, or 8192 for the small sizes (8192/seqlen batch members share one 8192-point unit)."""
        return self.plan(device).fft_size

    def forward_host(self, u, k, pregate=None, postgate=None, out=None, device=None):
        """Forward on HOST tensors: y = forward(u.cuda(), k.cuda(), ...).cpu(), with the host->device copies, the
        convolution and the device->host copy pipelined over batch chunks (bffc_fwd_host, include/bffc.h), so both
        PCIe directions are busy at once.  u / gates / out: (B, H, L) host tensors of the module dtype (pin them, or
        the copies serialise); k: (H, Lk) fp32, host or device.  The result is complete once the current CUDA stream
        of `device` is (the call is asynchronous).  Inference only (no autograd)."""
        device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        return _forward_host(self, u, k, pregate, postgate, out, device)

    def forward(self, u, k, pregate=None, postgate=None):
        if pregate is not None or postgate is not None:
            assert pregate is not None and postgate is not None       # conv.py:557-558
            return GatedFlashFFTConvFunc.apply(u, k, self, pregate, postgate)
        return FlashFFTConvFunc.apply(u, k, self)


def _forward_host(mod, u, k, pregate, postgate, out, device):
    """See FlashFFTConv.forward_host."""
    if (pregate is None) != (postgate is None):
        raise AssertionError('pregate and postgate must both be given or both be None')       # conv.py:557-558
    gates = [g for g in (pregate, postgate) if g is not None]
    for t in [u] + gates:
        if t.is_cuda or t.dtype != mod.dtype or t.dim() != 3 or not t.is_contiguous() or t.shape != u.shape:
            raise RuntimeError(f'forward_host: u / gates must be contiguous host (B, H, L) tensors of dtype {mod.dtype}')
    B, H, L = u.shape
    if k.dim() != 2 or k.shape[0] != H or k.shape[1] > mod.seqlen or L > mod.seqlen:
        raise RuntimeError(f'k must be (H={H}, Lk<={mod.seqlen}) and L <= seqlen, got {tuple(k.shape)}, L={L}')
    if L % mod.plan(device).length_multiple:
        raise RuntimeError(f'forward_host: L={L} must be a multiple of bffc_length_multiple(); pad on the host or '
                           'use forward() with device tensors')
    if out is None:
        out = torch.empty(u.shape, dtype=u.dtype, pin_memory=True)
    if out.is_cuda or out.shape != u.shape or out.dtype != u.dtype or not out.is_contiguous():
        raise RuntimeError('forward_host: out must be a contiguous host tensor like u')
    plan = mod.plan(device)
    with torch.cuda.device(device):
        kf_engine = _kf_engine_for(mod, plan, k if k.is_cuda else k.to(device, non_blocking=True), cache_key=k)
        nws = _lib.lib().bffc_host_workspace_bytes(plan.handle, B, H, L, 1 if gates else 0)
        ws = mod._host_ws.get((device, nws))
        if ws is None:
            mod._host_ws.clear()
            ws = mod._host_ws[(device, nws)] = torch.empty(nws, dtype=torch.uint8, device=device)
        rc = _lib.lib().bffc_fwd_host(plan.handle, _ptr(u), _ptr(kf_engine), _ptr(pregate), _ptr(postgate),
                                      _ptr(out), B, H, L, _ptr(ws), nws, _stream())
        if rc:
            torch.cuda.synchronize(device)     # nothing may still be copying into / out of buffers we are about to drop
            _lib.check(rc)
        mod.__dict__['last_launches'] = 1 + _lib.lib().bffc_last_launch_count()
        # the library joins its internal streams back into the current stream before returning, so the caching
        # allocator (stream-ordered on the current stream) may recycle kf_engine / ws after this point
    return out


def _check_inputs(u, k, mod, gates=()):
    if not u.is_cuda:
        raise RuntimeError('u must be a CUDA tensor (bffc has no CPU path)')          # monarch_fwd.h:7-13
    if u.dtype != mod.dtype:
        raise RuntimeError(f'u must have dtype {mod.dtype}, got {u.dtype}')
    if u.dim() != 3 or not u.is_contiguous():
        raise RuntimeError('u must be a contiguous (B, H, L) tensor')
    B, H, L = u.shape
    if k.dim() != 2 or k.shape[0] != H or k.shape[1] > mod.seqlen:
        raise RuntimeError(f'k must be (H={H}, Lk<={mod.seqlen}), got {tuple(k.shape)}')
    if L > mod.seqlen:
        raise RuntimeError(f'L={L} exceeds seqlen={mod.seqlen}')
    for g in gates:
        if g.shape != u.shape or g.dtype != u.dtype or not g.is_contiguous() or not g.is_cuda:
            raise RuntimeError('gates must match u in shape, dtype, device and be contiguous')
    return B, H, L


def _pack_kf_from_natural(mod, plan, k_f, conj):
    """rfft k_f -> engine-order packed (H, N) 4-byte complex, scaled 1/N (replaces conv.py:640)."""
    kf_engine = torch.empty((k_f.shape[0], mod.fft_size(k_f.device)), dtype=torch.int32, device=k_f.device)
    _lib.check(_lib.lib().bffc_kf_pack_rfft(plan.handle, _ptr(torch.view_as_real(k_f)), _ptr(kf_engine),
                                            int(k_f.shape[0]), int(conj), _stream()))
    return kf_engine


def _filter_workspace(plan, H, device):
    n = plan.filter_workspace_bytes(H)
    return (torch.empty(n, dtype=torch.uint8, device=device) if n else None), n


def _pack_kf(mod, plan, k, conj=0):
    """k (H, Lk) fp32 device -> engine-order packed spectrum (H, N) int32 words by the library's own fp32 FFT
    (bffc_kf_from_filter): one launch for engine size 8192, column + row FFT launches per L2-sized channel group for
    the composite sizes (replaces conv.py:575 + :640)."""
    k32 = k.detach()
    if k32.dtype != torch.float32 or not k32.is_contiguous():
        k32 = k32.to(torch.float32).contiguous()
    H, Lk = k32.shape
    kf_engine = torch.empty((H, plan.fft_size), dtype=torch.int32, device=k.device)
    ws, ws_bytes = _filter_workspace(plan, H, k.device)
    _lib.check(_lib.lib().bffc_kf_from_filter(plan.handle, _ptr(k32), int(Lk), _ptr(kf_engine), int(H), int(conj),
                                              _ptr(ws), ws_bytes, _stream()))
    mod.__dict__['last_launches'] = _lib.lib().bffc_last_launch_count()
    return kf_engine


def _kf_engine_for(mod, plan, k, cache_key=None):
    """Engine-order spectrum of `k`, cached in eval mode while the same tensor object is unmodified."""
    key = k if cache_key is None else cache_key
    use_cache = not mod.training
    if use_cache and mod._kf_cache is not None:
        ref, ver, dev, kf = mod._kf_cache
        if ref() is key and ver == key._version and dev == k.device:
            return kf
    kf = _pack_kf(mod, plan, k)
    mod.__dict__['_kf_cache'] = (weakref.ref(key), key._version, k.device, kf) if use_cache else None
    return kf


def _pad_len(mod, device, L):
    q = mod.plan(device).length_multiple
    return (L + q - 1) // q * q


def _padded(t, Lp):
    """zero-extend (B,H,L) to (B,H,Lp): identical operator (implicit zero padding), used for lengths the kernels'
    tiling does not take directly (the reference only requires L even, README.md:270)."""
    if t is None or t.shape[-1] == Lp:
        return t
    return torch.nn.functional.pad(t, (0, Lp - t.shape[-1]))          # one kernel: copy + zero tail


def _workspace(plan, B, H, L, gated, backward, device):
    n = plan.workspace_bytes(B, H, L, gated, backward)
    return (torch.empty(n, dtype=torch.uint8, device=device) if n else None), n


class _on_device:
    """torch.cuda.device(dev) only when dev is not already current (the context manager costs several microseconds)."""

    def __init__(self, device):
        self.ctx = None if torch.cuda.current_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _fwd(mod, u, k, pregate, postgate):
    L0 = u.shape[-1]
    Lp = _pad_len(mod, u.device, L0)
    if Lp != L0:
        y, kf = _fwd(mod, _padded(u, Lp), k, _padded(pregate, Lp), _padded(postgate, Lp))
        return y[..., :L0].contiguous(), kf
    B, H, L = u.shape
    plan = mod.plan(u.device)
    with _on_device(u.device):
        mod.__dict__['last_launches'] = 0
        kf_engine = _kf_engine_for(mod, plan, k)
        y = torch.empty_like(u)
        ws, ws_bytes = _workspace(plan, B, H, L, pregate is not None, False, u.device)
        _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(u), _ptr(kf_engine), _ptr(pregate), _ptr(postgate),
                                       _ptr(y), B, H, L, _ptr(ws), ws_bytes, _stream()))
        mod.__dict__['last_launches'] += _lib.lib().bffc_last_launch_count()
    return y, kf_engine


def _bwd(mod, dout, u, kf_engine, k_len, pregate, postgate):
    """du, dk[, dpregate, dpostgate] — reference: FlashFFTConvFunc.backward, conv.py:1737-1822."""
    L0 = u.shape[-1]
    Lp = _pad_len(mod, u.device, L0)
    if Lp != L0:
        r = _bwd(mod, _padded(dout, Lp), _padded(u, Lp), kf_engine, k_len, _padded(pregate, Lp), _padded(postgate, Lp))
        cut = lambda t: None if t is None else t[..., :L0].contiguous()
        return cut(r[0]), r[1], cut(r[2]), cut(r[3])
    B, H, L = u.shape
    N = mod.fft_size(u.device)
    plan = mod.plan(u.device)
    dout = dout.contiguous()                                          # conv.py:1742
    with _on_device(u.device):
        du = torch.empty_like(u)
        dkf_engine = torch.empty((H, N, 2), dtype=torch.float32, device=u.device)
        dpre = torch.empty_like(u) if pregate is not None else None
        dpost = torch.empty_like(u) if pregate is not None else None
        ws, ws_bytes = _workspace(plan, B, H, L, pregate is not None, True, u.device)
        # kf_engine_conj = NULL: the kernels conjugate the forward's spectrum in their pointwise multiply
        _lib.check(_lib.lib().bffc_bwd(plan.handle, _ptr(dout), _ptr(u), _ptr(kf_engine), None, _ptr(pregate),
                                       _ptr(postgate), _ptr(du), _ptr(dkf_engine), _ptr(dpre), _ptr(dpost),
                                       B, H, L, _ptr(ws), ws_bytes, _stream()))
        mod.__dict__['last_launches'] = _lib.lib().bffc_last_launch_count()
        # the kernels accumulate unnormalised pair-packed spectra in engine order; the reference takes
        # ifft(dk_f).real[..., :k_len] (conv.py:1817-1820): inverse fp32 FFT straight from engine order, 1/N, real part
        # (only the Hermitian part of dk_f contributes), sum over the batch-member blocks of the small sizes, [:k_len]
        dk = torch.empty((H, k_len), dtype=torch.float32, device=u.device)
        fws, fws_bytes = _filter_workspace(plan, H, u.device)
        _lib.check(_lib.lib().bffc_dk_from_dkf(plan.handle, _ptr(dkf_engine), _ptr(dk), int(k_len), H, _ptr(fws), fws_bytes,
                                               _stream()))
        mod.__dict__['last_launches'] += _lib.lib().bffc_last_launch_count()
    return du, dk, dpre, dpost


class FlashFFTConvFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, k, mod):
        _check_inputs(u, k, mod)
        y, kf_engine = _fwd(mod, u, k, None, None)
        ctx.mod = mod
        ctx.k_len = k.shape[-1]
        if mod.training:                                              # conv.py:587-588
            ctx.save_for_backward(u, kf_engine)
        return y

    @staticmethod
    def backward(ctx, dout):
        u, kf_engine = ctx.saved_tensors
        du, dk, _, _ = _bwd(ctx.mod, dout, u, kf_engine, ctx.k_len, None, None)
        return du, dk, None                                           # conv.py:1822


class GatedFlashFFTConvFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, k, mod, pregate, postgate):
        _check_inputs(u, k, mod, (pregate, postgate))
        y, kf_engine = _fwd(mod, u, k, pregate, postgate)
        ctx.mod = mod
        ctx.k_len = k.shape[-1]
        if mod.training:
            ctx.save_for_backward(u, kf_engine, pregate, postgate)
        return y

    @staticmethod
    def backward(ctx, dout):
        u, kf_engine, pregate, postgate = ctx.saved_tensors
        du, dk, dpre, dpost = _bwd(ctx.mod, dout, u, kf_engine, ctx.k_len, pregate, postgate)
        return du, dk, None, dpre, dpost                              # conv.py:3939

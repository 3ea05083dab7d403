"""Host-side mirror of the reference operator API for the fused FFT-convolution path.

Drop-in for `flashfftconv.FlashFFTConv` (reference flashfftconv/conv.py:71-560): same constructor
`FlashFFTConv(seqlen, dtype=torch.float16, use_32_butterfly=True)`, same
`forward(u, k, pregate=None, postgate=None)`, same autograd contract
(`backward -> (du, dk, None[, dpregate, dpostgate])`, conv.py:1822, :3939).

All arithmetic on the hot path happens in libbffc.so (hand-written sm_100a CUDA, C ABI in
include/bffc.h).  PyTorch is used for device memory, streams and — exactly as the reference does at
conv.py:575 and :1817 — for the fp32 FFT of the filter `k` and the inverse FFT of `dk_f`.
"""
import ctypes

import torch

from . import _lib

_DT = {torch.bfloat16: _lib.BFFC_DTYPE_BF16, torch.float16: _lib.BFFC_DTYPE_FP16}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Plan:
    """Owns one bffc_plan per (seqlen, dtype, device)."""

    def __init__(self, seqlen, dtype, device):
        self.handle = ctypes.c_void_p(0)
        self.device = device
        with torch.cuda.device(device):
            _lib.check(_lib.lib().bffc_plan_create(ctypes.byref(self.handle), int(seqlen), _DT[dtype]))

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
        self._plans = {}

    def plan(self, device):
        key = (device.type, device.index)
        if key not in self._plans:
            self._plans[key] = _Plan(self.seqlen, self.dtype, device)
        return self._plans[key]

    def forward(self, u, k, pregate=None, postgate=None):
        if pregate is not None or postgate is not None:
            assert pregate is not None and postgate is not None       # conv.py:557-558
            return GatedFlashFFTConvFunc.apply(u, k, self, pregate, postgate)
        return FlashFFTConvFunc.apply(u, k, self)


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


def _pack_kf(mod, plan, k, conj):
    """k (H, Lk) fp32 -> engine-order packed k_f (H, N) 4-byte complex; reference: conv.py:575 + :640."""
    N = mod.seqlen
    k_f = torch.fft.fft(k.to(torch.float32), n=N).contiguous()        # complex64, natural order
    kf_engine = torch.empty((k.shape[0], N), dtype=torch.int32, device=k.device)
    _lib.check(_lib.lib().bffc_kf_pack(plan.handle, _ptr(torch.view_as_real(k_f)), _ptr(kf_engine),
                                       int(k.shape[0]), int(conj), _stream()))
    return kf_engine


def _fwd(mod, u, k, pregate, postgate):
    B, H, L = u.shape
    plan = mod.plan(u.device)
    with torch.cuda.device(u.device):
        kf_engine = _pack_kf(mod, plan, k, conj=0)
        y = torch.empty_like(u)
        ws_bytes = _lib.lib().bffc_workspace_bytes(plan.handle, B, H, L)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=u.device) if ws_bytes else None
        _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(u), _ptr(kf_engine), _ptr(pregate), _ptr(postgate),
                                       _ptr(y), B, H, L, _ptr(ws), ws_bytes, _stream()))
    return y, kf_engine


class FlashFFTConvFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, k, mod):
        _check_inputs(u, k, mod)
        y, kf_engine = _fwd(mod, u, k, None, None)
        ctx.mod = mod
        ctx.k_len = k.shape[-1]
        if mod.training:                                              # conv.py:587-588
            ctx.save_for_backward(u, k)
        return y

    @staticmethod
    def backward(ctx, dout):
        raise NotImplementedError('bffc backward is not implemented yet')


class GatedFlashFFTConvFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, k, mod, pregate, postgate):
        _check_inputs(u, k, mod, (pregate, postgate))
        y, kf_engine = _fwd(mod, u, k, pregate, postgate)
        ctx.mod = mod
        ctx.k_len = k.shape[-1]
        if mod.training:
            ctx.save_for_backward(u, k, pregate, postgate)
        return y

    @staticmethod
    def backward(ctx, dout):
        raise NotImplementedError('bffc backward is not implemented yet')

"""Partial and frequency-sparse convolutions through the fused engine.

The reference ships these two operators as plain-PyTorch examples (flashfftconv/sparse_conv.py:9-38: `PartialFFTConv`
truncates the filter to its first N_partial taps, `FrequencySparseFFTConv` zeroes the rfft bins from N_partial // 2 up;
both convolve at FFT size N = 2 L and keep the first L outputs).  Same classes, same call `forward(x, k)`, but the
convolution itself is one `FlashFFTConv(2 L)` launch sequence on the 16-bit engine:

* partial: the engine takes filters shorter than the sequence natively (`k: (H, Lk <= seqlen)`, zero-extended inside the
  filter-side FFT kernel), so truncation is a view;
* frequency-sparse: the masked half-spectrum goes through `bffc_kf_pack_rfft` (Hermitian completion, engine order, 1/N,
  cast) — the one place this package computes an FFT outside the library, because the caller's operator is defined on
  `torch.fft.rfft` bins.  Forward only (the reference example has no custom backward either; use it for inference).
"""
import torch

from . import _lib
from .conv import FlashFFTConv, _pack_kf_from_natural, _ptr, _stream, _workspace, _check_inputs, _on_device


class _EngineCache(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self._convs = {}

    def conv(self, seqlen, dtype, device):
        key = (seqlen, dtype, str(device))
        if key not in self._convs:
            self._convs[key] = FlashFFTConv(seqlen, dtype=dtype).to(device)
        return self._convs[key]


class PartialFFTConv(_EngineCache):
    """y = (x * k[..., :N_partial])[..., :L], linear convolution (reference sparse_conv.py:9-23)."""

    def __init__(self, N_partial):
        super().__init__()
        self.N_partial = N_partial

    def forward(self, x, k):
        L = x.shape[-1]
        return self.conv(2 * L, x.dtype, x.device)(x, k[..., : self.N_partial].contiguous())


class FrequencySparseFFTConv(_EngineCache):
    """y = irfft(rfft(x, 2L) * mask(rfft(k, 2L)))[..., :L] with the bins from N_partial // 2 up zeroed
    (reference sparse_conv.py:25-38)."""

    def __init__(self, N_partial):
        super().__init__()
        self.N_partial = N_partial

    @torch.no_grad()
    def forward(self, x, k):
        B, H, L = x.shape
        mod = self.conv(2 * L, x.dtype, x.device)
        _check_inputs(x, k, mod)
        plan = mod.plan(x.device)
        if L % plan.length_multiple:
            raise RuntimeError(f'L={L} must be a multiple of {plan.length_multiple} for seqlen {2 * L}')
        with _on_device(x.device):
            # the caller's bins are those of an rfft at 2L; the small sizes live on the 8192-point grid, where bin j of 2L
            # is bin j * q (the pack kernel samples exactly those), so the cut-off scales by q
            n = plan.fft_size
            q = n // (2 * L)
            k_f = torch.fft.rfft(k.float(), n=n)
            k_f[..., (self.N_partial // 2) * q:] = 0
            kf_engine = _pack_kf_from_natural(mod, plan, k_f.contiguous(), 0)
            y = torch.empty_like(x)
            ws, ws_bytes = _workspace(plan, B, H, L, False, False, x.device)
            _lib.check(_lib.lib().bffc_fwd(plan.handle, _ptr(x), _ptr(kf_engine), None, None, _ptr(y), B, H, L,
                                           _ptr(ws), ws_bytes, _stream()))
        return y

"""Executable float64 model of the sm_100a kernel's dataflow for N = 128 x 64 (fwd3_r128.cuh).

Test infrastructure only.  It mirrors, stage by stage, the TMEM images the kernel produces
(lane = row, 128 fp32 columns; round 1 compared them with stage dumps of a bring-up build of the kernel, agreement
3e-7), and it proves (against numpy.fft) that the factorisation, the folded twiddles, the block
layouts and the k_f "engine order" are right before any GPU time is spent.

`quant=True` rounds every tensor-core operand to bf16 exactly where the kernel does, which gives the
expected rel-L2 error of the real kernel.
"""
import numpy as np

N = 8192
R = 128
M = 64


def bf16_round(x):
    """Round float64/32 array to bf16 (round-to-nearest-even), return float64."""
    f = np.asarray(x, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    u = ((u + r) >> 16) << 16
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def half_round(x):
    return np.asarray(x, dtype=np.float16).astype(np.float64)


def model_fwd(x0, x1, kf_nat, quant=False, ksteps=8):
    """x0, x1: real sequences (length L <= N, zero padded here); kf_nat: FFT_N(k) natural order (complex).
    Returns (y0, y1, stages) with stages = list of four (128,128) float64 TMEM images
    (cols [0,64) real part, [64,128) imaginary part): D1 outer DFT, D2 spectrum, D3 after inverse radix-64, D4."""
    q = bf16_round if quant else (lambda v: np.asarray(v, dtype=np.float64))
    qh = half_round if quant else (lambda v: np.asarray(v, dtype=np.float64))
    xr = np.zeros(N); xr[: len(x0)] = x0
    xi = np.zeros(N); xi[: len(x1)] = x1
    Xr = q(xr.reshape(R, M)); Xi = q(xi.reshape(R, M))       # tiles [i][j]
    nrows = 16 * ksteps
    mk = np.arange(R)
    ang = 2 * np.pi * ((mk[:, None] * mk[None, :]) % 128) / 128.0
    C = q(np.cos(ang)); S = q(np.sin(ang))
    stages = []
    # stage 1: F = C - iS
    Dre = C[:, :nrows] @ Xr[:nrows] + S[:, :nrows] @ Xi[:nrows]
    Dim = C[:, :nrows] @ Xi[:nrows] - S[:, :nrows] @ Xr[:nrows]
    stages.append(np.concatenate([Dre, Dim], axis=1))
    Y = Dre + 1j * Dim                                         # [k1][j]
    k1 = np.arange(R)[:, None]
    j = np.arange(M)[None, :]
    tw = np.exp(-2j * np.pi * ((k1 * j) % N) / N)
    tw = qh(tw.real) + 1j * qh(tw.imag)                        # kernel keeps the twiddles as half2
    # pass 1
    Y1 = Y * tw
    Y1 = q(Y1.real) + 1j * q(Y1.imag)
    # stage 2: radix-64, G[j,k2] = exp(-2 pi i j k2 / 64)
    e = np.arange(M)
    angg = -2 * np.pi * ((e[:, None] * e[None, :]) % 64) / 64.0
    G = q(np.cos(angg)) + 1j * q(np.sin(angg))
    Z = Y1 @ G                                                  # [k1][k2]
    stages.append(np.concatenate([Z.real, Z.imag], axis=1))
    # pass 3: * k_f[k1 + 128*k2] / N, bf16
    kfe = (np.asarray(kf_nat).reshape(M, R).T) / N             # [k1][k2]
    kfe = q(kfe.real) + 1j * q(kfe.imag)
    V = Z * kfe
    V = q(V.real) + 1j * q(V.imag)
    # stage 3: inverse radix-64
    Yi = V @ np.conj(G)                                         # [k1][j]
    stages.append(np.concatenate([Yi.real, Yi.imag], axis=1))
    # pass 5: * conj tw -> smem rows
    Yn = Yi * np.conj(tw)
    Yr = q(Yn.real); Yim = q(Yn.imag)
    # stage 4: conj F = C + iS
    Ore = C @ Yr - S @ Yim
    Oim = C @ Yim + S @ Yr
    stages.append(np.concatenate([Ore, Oim], axis=1))
    return Ore.reshape(-1), Oim.reshape(-1), stages


def ref_conv(x, k, n=N):
    """float64 statement of tests/test_flashfftconv.py:5-13 (circular conv mod n, truncated)."""
    L = len(x)
    return np.fft.ifft(np.fft.fft(x, n) * np.fft.fft(k, n)).real[:L]


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    x0 = rng.standard_normal(N); x1 = rng.standard_normal(N)
    k = rng.standard_normal(N) / np.sqrt(N)
    kf = np.fft.fft(k, N)
    y0, y1, st = model_fwd(x0, x1, kf)
    r0, r1 = ref_conv(x0, k), ref_conv(x1, k)
    print('exact model max err', np.abs(y0 - r0).max(), np.abs(y1 - r1).max())
    y0q, y1q, _ = model_fwd(bf16_round(x0), bf16_round(x1), kf, quant=True)
    r0q = ref_conv(bf16_round(x0), k)
    print('bf16 model rel-L2', np.linalg.norm(y0q - r0q) / np.linalg.norm(r0q))

"""Executable float64 model of the sm_100a kernel's dataflow for N = 128 x 64 (fwd_r128.cuh).

Test infrastructure only.  It mirrors, stage by stage, the TMEM images the kernel produces
(lane = row, 128 fp32 columns) so the GPU stage dumps from `bffc_debug_fwd_stages` can be compared
with it, and it proves (against numpy.fft) that the factorisation, the folded twiddles, the block
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


def engine_perm():
    """engine index e = k1*64 + 8a + d  ->  natural frequency k = k1 + 128*(a + 8d)."""
    k1 = np.arange(128)[:, None, None]
    a = np.arange(8)[None, :, None]
    d = np.arange(8)[None, None, :]
    return (k1 + 128 * (a + 8 * d)).reshape(-1)


def model_fwd(x0, x1, kf_nat, quant=False, ksteps=8):
    """x0, x1: real sequences (length L <= N, zero padded here); kf_nat: FFT_N(k) natural order (complex).
    Returns (y0, y1, stages) with stages = list of six (128,128) float64 images."""
    q = bf16_round if quant else (lambda v: np.asarray(v, dtype=np.float64))
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
    j1, j2 = j // 8, j % 8
    twA = np.exp(-2j * np.pi * (k1 * 8 * j1) / N)
    twB = np.exp(-2j * np.pi * (k1 * j2) / N)
    # pass 1
    Y1 = Y * twA
    Y1 = q(Y1.real) + 1j * q(Y1.imag)
    Y1 = Y1.reshape(R, 8, 8)                                   # [k1][j1][j2]
    # stage 2a: contract j1 with F8[j1,a]
    e8 = np.arange(8)
    F8 = np.exp(-2j * np.pi * (e8[:, None] * e8[None, :]) / 8)
    F8q = q(F8.real) + 1j * q(F8.imag)
    U = np.einsum('kjt,ja->kta', Y1, F8q)                      # [k1][j2][a]
    stages.append(np.concatenate([U.real, U.imag], axis=2).reshape(R, 128))   # block j2: [re a | im a]
    # pass 2: * twB[j2]; regroup [k1][a][j2]
    U2 = U * np.exp(-2j * np.pi * (np.arange(R)[:, None, None] * e8[None, :, None]) / N)
    U2 = q(U2.real) + 1j * q(U2.imag)
    U2 = U2.transpose(0, 2, 1)                                 # [k1][a][j2]
    # stage 2b: block a: G_a[j2,d] = exp(-2 pi i (a j2/64 + j2 d/8))
    a_ = e8[:, None, None]; j2_ = e8[None, :, None]; d_ = e8[None, None, :]
    G = np.exp(-2j * np.pi * (a_ * j2_ / 64.0 + j2_ * d_ / 8.0))           # [a][j2][d]
    Gq = q(G.real) + 1j * q(G.imag)
    Z = np.einsum('kat,atd->kad', U2, Gq)                      # [k1][a][d]
    stages.append(np.concatenate([Z.real, Z.imag], axis=2).reshape(R, 128))   # block a: [re d | im d]
    # pass 3: * k_f (engine order), scaled 1/N, bf16
    kfe = (np.asarray(kf_nat)[engine_perm()] / N).reshape(R, 8, 8)
    kfe = q(kfe.real) + 1j * q(kfe.imag)
    V = Z * kfe
    V = q(V.real) + 1j * q(V.imag)
    # stage 3b: H_a[d,j2] = conj(G_a[j2,d])
    Hq = np.conj(Gq).transpose(0, 2, 1)                        # [a][d][j2]
    Ui = np.einsum('kad,adt->kat', V, Hq)                      # [k1][a][j2]
    stages.append(np.concatenate([Ui.real, Ui.imag], axis=2).reshape(R, 128))  # block a: [re j2 | im j2]
    # pass 4: * conj twB[j2]; regroup [k1][j2][a]
    Ui = Ui * np.exp(2j * np.pi * (np.arange(R)[:, None, None] * e8[None, None, :]) / N)
    Ui = q(Ui.real) + 1j * q(Ui.imag)
    Ui = Ui.transpose(0, 2, 1)                                 # [k1][j2][a]
    # stage 3a: iF8[a,j1]
    Yi = np.einsum('kta,aj->ktj', Ui, np.conj(F8q))            # [k1][j2][j1]
    stages.append(np.concatenate([Yi.real, Yi.imag], axis=2).reshape(R, 128))  # block j2: [re j1 | im j1]
    # pass 5: * conj twA[j1] -> smem rows [k1][j = 8 j1 + j2]
    Yn = Yi.transpose(0, 2, 1).reshape(R, M) * np.conj(twA)
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

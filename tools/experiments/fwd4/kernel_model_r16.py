"""Executable model of the flop-lean fused kernel's dataflow, N = 8192 = 16 x 16 x 32 (fwd4_r16.cuh).

Test infrastructure only.  It mirrors the kernel at the level the bugs live at: the byte images the passes write to
shared memory (128-byte swizzle, MN-major / K-major operand tiles), the way the tensor core reads them back through the
canonical UMMA layouts named by the shared-memory descriptors, the TMEM images (lane, column) after every stage, the
folded chunk twiddles in the DFT operands, the k_f "engine order v2" and the transposing passes of the inverse chain.
Checked against numpy.fft in tests/test_oracle.py; the GPU stage dumps are compared with `stages`.

Index algebra (n = time index of the pair-packed complex sequence z = x_b + i x_{b+1}, f = frequency):
  forward, decimation in frequency, radices 16 (a), 16 (b), 32 (c):
     n = 512 a + n'',  n'' = 128 t + m = 32 b + c,  m = 64 h + e (S1 lane),  c = 8 c_hi + c_lo,  b = 4 t + (m >> 5)
     f = q1 + 16 q2 + 256 q3
  inverse, decimation in frequency on the spectrum, radices 32 (q3 -> c), 16 (q2 -> b), 16 (q1 -> a).
Stage s writes TMEM[lane][col]; the following pass reads it (thread = lane, or the 16x256b fragment layout for the two
transposing passes), multiplies by the twiddle and writes the next A operand.
"""
import numpy as np

N = 8192
TILE = 16384          # bytes of one plane (real or imaginary parts, 16-bit elements)


def bf16_round(x):
    f = np.asarray(x, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    u = ((u + r) >> 16) << 16
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def W(n, e):
    return np.exp(-2j * np.pi * (np.asarray(e) % n) / n)


# ----------------------------------------------------------------------------- shared-memory byte addressing
def swz(addr):
    """128-byte swizzle: 16-byte chunk index (bits 4-6) XOR row-in-atom (bits 7-9); tiles are 1024-byte aligned."""
    return addr ^ (((addr >> 7) & 7) << 4)


def row_elem_addr(row, elem):
    """element `elem` (0..63) of 128-byte row `row` of a plane, swizzled"""
    return swz(row * 128 + elem * 2)


class Plane:
    """one 16 KB operand plane as 8192 16-bit elements addressed by byte offset"""

    def __init__(self):
        self.v = np.full(TILE // 2, np.nan)

    def st(self, addr, val):
        assert addr % 2 == 0 and 0 <= addr < TILE
        self.v[addr // 2] = val

    def ld(self, addr):
        return self.v[addr // 2]


def umma_read_mn(plane, start, lbo, k_rows, k0=0):
    """A[m][k] (m < 128, k < k_rows) of an MN-major, 128B-swizzled operand: canonical layout
    ((8 chunks x 8 elems, atoms),(8 rows, groups)) : ((1, LBO),(128 B, SBO = 1024)); a K step of 16 rows = +2048 B"""
    A = np.empty((128, k_rows))
    for m in range(128):
        for k in range(k_rows):
            kk = k0 + k
            addr = start + (m // 64) * lbo + (kk // 8) * 1024 + (kk % 8) * 128 + (m % 64) * 2
            A[m, k] = plane.ld(swz(addr))
    return A


def umma_read_k(plane, byte_off, k_cols):
    """A[m][k] of a K-major 128B-swizzled tile [128 rows][64 elems]: row m at 128 m, K elements from byte_off"""
    A = np.empty((128, k_cols))
    for m in range(128):
        for k in range(k_cols):
            A[m, k] = plane.ld(swz(m * 128 + byte_off + 2 * k))
    return A


# ----------------------------------------------------------------------------- DFT operands (B matrices)
def b_pair(F):
    """complex r x r' matrix -> (Br, Bi): real [K = r][N = 2 r'] operands for the real / imaginary input plane;
    output columns [re | im]."""
    return np.concatenate([F.real, F.imag], axis=1), np.concatenate([-F.imag, F.real], axis=1)


def dft(r, sign):
    k = np.arange(r)
    return np.exp(sign * 2j * np.pi * np.outer(k, k) / r)


def b_radix16_fwd(chunk, q):
    """forward radix-16 operand of chunk `chunk`: W16^{a q} * W_64^{q chunk} (the part of the following twiddle that
    depends on (output digit, chunk) only is folded into the matrix; S1 and S2 share these four matrices)"""
    F = dft(16, -1) * W(64, np.arange(16) * chunk)[None, :]
    Br, Bi = b_pair(F)
    return q(Br), q(Bi)


# ----------------------------------------------------------------------------- k_f engine order v2
def kf_engine_freq(v, lane, j):
    """frequency held by entry j (0..3) of 16-byte vector v (0..15) of lane `lane`: pass 3's thread (lane m3) reads
    vector v = 8 q1_hi + g for columns q3 = 4 g + j of chunk q1_hi"""
    q1_hi, g = v >> 3, v & 7
    q2_hi, q1_lo, q2_lo = lane >> 6, (lane >> 3) & 7, lane & 7
    return (8 * q1_hi + q1_lo) + 16 * (8 * q2_hi + q2_lo) + 256 * (4 * g + j)


def model_fwd4(x0, x1, kf_nat, quant=False, L=None):
    """x0, x1: real sequences (zero padded to N here); kf_nat: FFT_N(k), natural order.  Returns (y0, y1, stages):
    stages = six (128, 128) TMEM images D1, D2, D3, D3', D2', D1' (columns as the kernel lays them out)."""
    q = bf16_round if quant else (lambda v: np.asarray(v, dtype=np.float64))
    xr = np.zeros(N); xr[: len(x0)] = x0
    xi = np.zeros(N); xi[: len(x1)] = x1
    stages = []

    # ---------------- TMA load: 5-D map (e, a, pair, th, channel), smem row = 16 * th + a, 128B swizzle
    Pr, Pi = Plane(), Plane()
    for n in range(N):
        a, th, e = n >> 9, (n >> 6) & 7, n & 63
        Pr.st(row_elem_addr(16 * th + a, e), q(xr[n]))
        Pi.st(row_elem_addr(16 * th + a, e), q(xi[n]))

    # ---------------- S1: chunk t, A = rows (t, h, a) MN-major (LBO 2048 between the two 64-element atoms)
    D = np.zeros((128, 128))
    for t in range(4):
        Br, Bi = b_radix16_fwd(t, q)
        Ar, Ai = umma_read_mn(Pr, 4096 * t, 2048, 16), umma_read_mn(Pi, 4096 * t, 2048, 16)
        D[:, 32 * t: 32 * t + 32] = Ar @ Br + Ai @ Bi
    stages.append(D.copy())

    # ---------------- P1 (thread = lane m): * W_8192^{q1 m}; A2: chunk c_hi, atom q1_hi, row b, element 8 c_lo + q1_lo
    Pr, Pi = Plane(), Plane()
    for m in range(128):
        c, m_hi = m & 31, m >> 5
        c_hi, c_lo = c >> 3, c & 7
        for t in range(4):
            b = 4 * t + m_hi
            for q1 in range(16):
                val = (D[m, 32 * t + q1] + 1j * D[m, 32 * t + 16 + q1]) * W(N, q1 * m)
                row = (2 * c_hi + (q1 >> 3)) * 16 + b
                Pr.st(row_elem_addr(row, 8 * c_lo + (q1 & 7)), q(val.real))
                Pi.st(row_elem_addr(row, 8 * c_lo + (q1 & 7)), q(val.imag))
    # ---------------- S2: chunk c_hi; D2 lane = 64 q1_hi + 8 c_lo + q1_lo, cols 32 c_hi + 16 ri + q2
    D = np.zeros((128, 128))
    for ch in range(4):
        Br, Bi = b_radix16_fwd(ch, q)
        Ar, Ai = umma_read_mn(Pr, 4096 * ch, 2048, 16), umma_read_mn(Pi, 4096 * ch, 2048, 16)
        D[:, 32 * ch: 32 * ch + 32] = Ar @ Br + Ai @ Bi
    stages.append(D.copy())

    # ---------------- P2 (thread = lane): * W_512^{q2 c_lo}; A3: chunk q1_hi, atom q2_hi, row c, element 8 q1_lo + q2_lo
    Pr, Pi = Plane(), Plane()
    for m in range(128):
        q1_hi, c_lo, q1_lo = m >> 6, (m >> 3) & 7, m & 7
        for ch in range(4):
            c = 8 * ch + c_lo
            for q2 in range(16):
                val = (D[m, 32 * ch + q2] + 1j * D[m, 32 * ch + 16 + q2]) * W(512, q2 * c_lo)
                row = (2 * q1_hi + (q2 >> 3)) * 32 + c
                Pr.st(row_elem_addr(row, 8 * q1_lo + (q2 & 7)), q(val.real))
                Pi.st(row_elem_addr(row, 8 * q1_lo + (q2 & 7)), q(val.imag))
    # ---------------- S3: radix 32 over c; chunk q1_hi (start 8192 t, LBO 4096, two K steps);
    #                  D3 lane = 64 q2_hi + 8 q1_lo + q2_lo, cols 64 q1_hi + 32 ri + q3
    F32r, F32i = b_pair(dft(32, -1))
    F32r, F32i = q(F32r), q(F32i)
    D = np.zeros((128, 128))
    for t in range(2):
        Ar, Ai = umma_read_mn(Pr, 8192 * t, 4096, 32), umma_read_mn(Pi, 8192 * t, 4096, 32)
        D[:, 64 * t: 64 * t + 64] = Ar @ F32r + Ai @ F32i
    stages.append(D.copy())

    # ---------------- P3 (thread = lane): * k_f (engine order v2, 1/N folded in); K-major tile, row = lane, element 32 q1_hi + q3
    kfe = np.empty((16, 128, 4), dtype=complex)
    for v in range(16):
        for lane in range(128):
            for j in range(4):
                kfe[v, lane, j] = kf_nat[kf_engine_freq(v, lane, j)] / N
    kfe = q(kfe.real) + 1j * q(kfe.imag)
    Pr, Pi = Plane(), Plane()
    for m in range(128):
        for q1_hi in range(2):
            for q3 in range(32):
                val = (D[m, 64 * q1_hi + q3] + 1j * D[m, 64 * q1_hi + 32 + q3]) * kfe[8 * q1_hi + (q3 >> 2), m, q3 & 3]
                Pr.st(row_elem_addr(m, 32 * q1_hi + q3), q(val.real))
                Pi.st(row_elem_addr(m, 32 * q1_hi + q3), q(val.imag))
    # ---------------- S3': inverse radix 32 over q3 (row local, K-major A); cols 64 q1_hi + 32 ri + c
    G32r, G32i = b_pair(dft(32, +1))
    G32r, G32i = q(G32r), q(G32i)
    D = np.zeros((128, 128))
    for t in range(2):
        Ar, Ai = umma_read_k(Pr, 64 * t, 32), umma_read_k(Pi, 64 * t, 32)
        D[:, 64 * t: 64 * t + 64] = Ar @ G32r + Ai @ G32i
    stages.append(D.copy())

    # ---------------- P4 (thread = lane (q2, q1_lo)): * conj W_8192^{c (q1 + 16 q2)}; A2': chunk c_hi, atom q1_hi, row q2,
    #                  element 8 q1_lo + c_lo
    Pr, Pi = Plane(), Plane()
    for m in range(128):
        q2, q1_lo = 8 * (m >> 6) + (m & 7), (m >> 3) & 7
        for q1_hi in range(2):
            q1 = 8 * q1_hi + q1_lo
            for c in range(32):
                val = (D[m, 64 * q1_hi + c] + 1j * D[m, 64 * q1_hi + 32 + c]) * np.conj(W(N, c * (q1 + 16 * q2)))
                row = (2 * (c >> 3) + q1_hi) * 16 + q2
                Pr.st(row_elem_addr(row, 8 * q1_lo + (c & 7)), q(val.real))
                Pi.st(row_elem_addr(row, 8 * q1_lo + (c & 7)), q(val.imag))
    # ---------------- S2': inverse radix 16 over q2; D2' lane = 64 q1_hi + 8 q1_lo + c_lo, cols 32 c_hi + 16 ri + b
    G16r, G16i = b_pair(dft(16, +1))
    G16r, G16i = q(G16r), q(G16i)
    D = np.zeros((128, 128))
    for ch in range(4):
        Ar, Ai = umma_read_mn(Pr, 4096 * ch, 2048, 16), umma_read_mn(Pi, 4096 * ch, 2048, 16)
        D[:, 32 * ch: 32 * ch + 32] = Ar @ G16r + Ai @ G16i
    stages.append(D.copy())

    # ---------------- P5 (TRANSPOSING: 16x256b fragments + stmatrix.trans): * conj W_256^{b q1};
    #                  A1': chunk c_hi, atom b_hi, row q1, element 8 b_lo + c_lo  (8 lanes c_lo = one 16-byte row of a matrix)
    Pr, Pi = Plane(), Plane()
    for m in range(128):
        q1, c_lo = 8 * (m >> 6) + ((m >> 3) & 7), m & 7
        for ch in range(4):
            for b in range(16):
                val = (D[m, 32 * ch + b] + 1j * D[m, 32 * ch + 16 + b]) * np.conj(W(256, b * q1))
                row = (2 * ch + (b >> 3)) * 16 + q1
                Pr.st(row_elem_addr(row, 8 * (b & 7) + c_lo), q(val.real))
                Pi.st(row_elem_addr(row, 8 * (b & 7) + c_lo), q(val.imag))
    # ---------------- S1': inverse radix 16 over q1; D1' lane = 64 b_hi + 8 b_lo + c_lo, cols 32 c_hi + 16 ri + a
    D = np.zeros((128, 128))
    for ch in range(4):
        Ar, Ai = umma_read_mn(Pr, 4096 * ch, 2048, 16), umma_read_mn(Pi, 4096 * ch, 2048, 16)
        D[:, 32 * ch: 32 * ch + 32] = Ar @ G16r + Ai @ G16i
    stages.append(D.copy())

    # ---------------- P6 (TRANSPOSING): 16-bit output tiles in the same 5-D TMA layout as the input
    Pr, Pi = Plane(), Plane()
    for m in range(128):
        b, c_lo = 8 * (m >> 6) + ((m >> 3) & 7), m & 7
        for ch in range(4):
            for a in range(16):
                n = 512 * a + 32 * b + 8 * ch + c_lo
                th, e = (n >> 6) & 7, n & 63
                Pr.st(row_elem_addr(16 * th + a, e), q(D[m, 32 * ch + a]))
                Pi.st(row_elem_addr(16 * th + a, e), q(D[m, 32 * ch + 16 + a]))
    y0 = np.empty(N); y1 = np.empty(N)
    for n in range(N):
        a, th, e = n >> 9, (n >> 6) & 7, n & 63
        y0[n] = Pr.ld(row_elem_addr(16 * th + a, e))
        y1[n] = Pi.ld(row_elem_addr(16 * th + a, e))
    return y0, y1, stages


def ref_conv(x, k, n=N):
    L = len(x)
    return np.fft.ifft(np.fft.fft(x, n) * np.fft.fft(k, n)).real[:L]


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    x0 = rng.standard_normal(N); x1 = rng.standard_normal(N)
    k = rng.standard_normal(N) / np.sqrt(N)
    kf = np.fft.fft(k, N)
    y0, y1, st = model_fwd4(x0, x1, kf)
    r0, r1 = ref_conv(x0, k), ref_conv(x1, k)
    print('exact model max err', np.abs(y0 - r0).max(), np.abs(y1 - r1).max())
    y0q, y1q, _ = model_fwd4(bf16_round(x0), bf16_round(x1), kf, quant=True)
    r0q = ref_conv(bf16_round(x0), k)
    print('bf16 model rel-L2', np.linalg.norm(y0q - r0q) / np.linalg.norm(r0q))

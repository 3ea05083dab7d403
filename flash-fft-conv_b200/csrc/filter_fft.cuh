// filter_fft.cuh — fp32 8192-point FFTs of the FILTER side of the path, in one launch each, directly in engine order:
//
//   kf_from_filter_kernel : k (H, Lk) fp32  ->  k_f engine words (16-bit complex, scaled, optionally conjugated)
//                           replaces  torch.fft.rfft(k, n=N) + bffc_kf_pack_rfft   (reference: conv.py:572-575 + :640)
//   dk_from_dkf_kernel    : dk_f engine (H, 8192) fp32 complex  ->  dk (H, Lk) fp32
//                           replaces  bffc_dkf_unpack + torch.fft.ifft(...).real[..., :Lk]   (reference: conv.py:1817-1820)
//
// The filter is fp32 in the reference and stays fp32 here (CUDA cores, not the 16-bit tensor pipe): its spectrum
// multiplies every sequence, so its error is not averaged out.  One CTA holds the whole 8192-point complex FFT in
// shared memory: in-place decimation-in-frequency, radices 16 x 16 x 32 (the last as 2 x 16), twiddles from a 64 KB
// plan table that stays in L1.  Output frequency f = q1 + 16 q2 + 256 q3 ends at position 512 q1 + 32 q2 + q3;
// the engine-order pack / the time-domain read-out undo that permutation on the fly, so nothing goes through HBM
// between the FFT and the layout change.  Two real filters share one complex FFT (z = k_a + i k_b).
#pragma once
#include "ptx.cuh"

namespace bffc {
namespace ffft {

constexpr int kN = 8192;
constexpr int kThreads = 256;
// position -> shared-memory slot: +1 per 32 and +1 per 512 keeps the three access patterns (stride 1, stride 32 and
// stride 512 across the lanes of a warp) free of bank conflicts
DEVINL int slot(int p) { return p + (p >> 5) + (p >> 9); }
constexpr int kSlots = kN + kN / 32 + kN / 512;
constexpr int kSmemBytes = kSlots * 8;

DEVINL float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DEVINL float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
DEVINL float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// multiply by sign * i
template <int SIGN> DEVINL float2 mul_i(float2 a) { return SIGN > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// 4-point DFT, exponent sign SIGN (-1 forward): y_c = sum_a x_a e^{SIGN 2 pi i a c / 4}
template <int SIGN>
DEVINL void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 s02 = cadd(x0, x2), d02 = csub(x0, x2), s13 = cadd(x1, x3), d13 = mul_i<SIGN>(csub(x1, x3));
  x0 = cadd(s02, s13); x2 = csub(s02, s13);
  x1 = cadd(d02, d13); x3 = csub(d02, d13);
}

// 16-point DFT in registers: v[m] -> v[q] (natural order on both sides)
template <int SIGN>
DEVINL void dft16(float2 (&v)[16]) {
  // m = 4a + b: DFT over a for every b -> t_b[c] kept at v[4c + b]
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4<SIGN>(v[b], v[4 + b], v[8 + b], v[12 + b]);
  // twiddle W16^{b c}
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
  const float sg = float(SIGN);
  const float2 w1 = make_float2(c1, sg * s1), w2 = make_float2(r2, sg * r2), w3 = make_float2(s1, sg * c1);
  const float2 w6 = make_float2(-r2, sg * r2), w9 = make_float2(-c1, -sg * s1);
  v[4 * 1 + 1] = cmul(v[4 * 1 + 1], w1); v[4 * 1 + 2] = cmul(v[4 * 1 + 2], w2); v[4 * 1 + 3] = cmul(v[4 * 1 + 3], w3);
  v[4 * 2 + 1] = cmul(v[4 * 2 + 1], w2); v[4 * 2 + 2] = mul_i<SIGN>(v[4 * 2 + 2]); v[4 * 2 + 3] = cmul(v[4 * 2 + 3], w6);
  v[4 * 3 + 1] = cmul(v[4 * 3 + 1], w3); v[4 * 3 + 2] = cmul(v[4 * 3 + 2], w6); v[4 * 3 + 3] = cmul(v[4 * 3 + 3], w9);
  // q = c + 4d: DFT over b for every c; v[4c + b] -> X[c + 4d] left at v[4c + d]
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4<SIGN>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  // transpose 4 x 4 so that v[q] = X[q]
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = c + 1; d < 4; ++d) { const float2 t = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = t; }
}

// one radix-16 decimation-in-frequency butterfly on the slots s0 + ss*m; output q is multiplied by tw[tstep*q]
// (tw[t] = e^{-2 pi i t / 8192}, plan table; conjugated for the inverse transform)
template <int SIGN>
DEVINL void butterfly16(float2* buf, int s0, int ss, const float2* __restrict__ tw, int tstep) {
  float2 v[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) v[m] = buf[s0 + ss * m];
  // twiddles W^{tstep q}: four table reads (q = 1, 2, 4, 8 — independent, issued before the butterfly), the rest by
  // products of those (depth <= 3 multiplications, so fp32 round-off stays at a few ulp)
  float2 w1 = __ldg(tw + tstep), w2 = __ldg(tw + 2 * tstep), w4 = __ldg(tw + 4 * tstep), w8 = __ldg(tw + 8 * tstep);
  dft16<SIGN>(v);
  if (SIGN > 0) { w1.y = -w1.y; w2.y = -w2.y; w4.y = -w4.y; w8.y = -w8.y; }
  const float2 w3 = cmul(w2, w1), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w9 = cmul(w8, w1), w10 = cmul(w8, w2), w12 = cmul(w8, w4);
  const float2 w7 = cmul(w6, w1), w11 = cmul(w10, w1), w13 = cmul(w12, w1), w14 = cmul(w12, w2);
  const float2 w15 = cmul(w14, w1);
  const float2 w[16] = {w1, w1, w2, w3, w4, w5, w6, w7, w8, w9, w10, w11, w12, w13, w14, w15};
  buf[s0] = v[0];
#pragma unroll
  for (int q = 1; q < 16; ++q) buf[s0 + ss * q] = cmul(v[q], w[q]);
}

// cos / sin of 2 pi m / 32, m = 0..15
__device__ constexpr float kCos32[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                         0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f,
                                         0.0f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                                         -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
__device__ constexpr float kSin32[16] = {0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                                         0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                                         1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                         0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};

// In-place 8192-point DFT of buf (natural order in, frequency f at position 512 (f & 15) + 32 ((f >> 4) & 15) + (f >> 8)).
// Slot arithmetic: +512 positions = +529 slots, +32 positions inside a block of 512 = +33 slots.
template <int SIGN>
DEVINL void fft8192(float2* buf, int tid, const float2* __restrict__ tw) {
  // stage A: span 8192, stride 512; twiddle W_8192^{t q}
  for (int t = tid; t < 512; t += kThreads) butterfly16<SIGN>(buf, t + (t >> 5), 529, tw, t);
  __syncthreads();
  // stage B: span 512, stride 32 inside each block of 512; twiddle W_512^{j q} = W_8192^{16 j q}
  for (int t = tid; t < 512; t += kThreads) butterfly16<SIGN>(buf, 529 * (t >> 5) + (t & 31), 33, tw, 16 * (t & 31));
  __syncthreads();
  // stage C: span 32 inside each block of 32, as radix 2 followed by radix 16: task = (block, even / odd outputs);
  // even: (a + b) -> outputs 2q, odd: (a - b) W_32^m -> outputs 2q + 1 (written without a branch: the two tasks of a
  // block are neighbouring lanes)
  for (int t = tid; t < 512; t += kThreads) {
    const int blk = t >> 1, odd = t & 1;
    const int s0 = 33 * blk + (blk >> 4);
    const float sb = odd ? -1.0f : 1.0f;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float2 a = buf[s0 + m], b = buf[s0 + m + 16];
      const float2 d = make_float2(fmaf(sb, b.x, a.x), fmaf(sb, b.y, a.y));
      const float2 w = make_float2(odd ? kCos32[m] : 1.0f, odd ? float(SIGN) * kSin32[m] : 0.0f);   // W_32^m or 1
      v[m] = cmul(d, w);
    }
    __syncwarp();                       // both tasks of a block have read before either writes
    dft16<SIGN>(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) buf[s0 + 2 * q + odd] = v[q];
  }
  __syncthreads();
}

DEVINL int pos_of_freq(int f) { return 512 * (f & 15) + 32 * ((f >> 4) & 15) + (f >> 8); }

// grid = ceil(H / 2): channels 2*blockIdx.x (real part) and 2*blockIdx.x + 1 (imaginary part)
template <int kFmt>
__global__ void __launch_bounds__(kThreads, 3) kf_from_filter_kernel(const float* __restrict__ k, int Lk, uint4* __restrict__ kf_eng,
                                                                     int H, float scale, int conj, const float2* __restrict__ tw) {
  extern __shared__ float2 fbuf[];
  const int tid = threadIdx.x, ha = 2 * blockIdx.x, hb = ha + 1;
  const float* ka = k + size_t(ha) * Lk;
  const float* kb = k + size_t(hb < H ? hb : ha) * Lk;
  for (int n = tid; n < kN; n += kThreads)
    fbuf[slot(n)] = n < Lk ? make_float2(ka[n], hb < H ? kb[n] : 0.f) : make_float2(0.f, 0.f);
  __syncthreads();
  fft8192<-1>(fbuf, tid, tw);
  // K_a[f] = (Z[f] + conj Z[-f]) / 2,  K_b[f] = (Z[f] - conj Z[-f]) / (2i); engine vector v = c*128 + k1 holds
  // frequencies k1 + 128 (4c + j), j = 0..3, as (re01, im01, re23, im23)
  using NT = Num<kFmt>;
  const float sa = 0.5f * scale, sgn = conj ? -1.f : 1.f;
  for (int v = tid; v < kN / 4; v += kThreads) {
    const int c = v >> 7, k1 = v & 127;
    float2 A[4], Bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = k1 + 128 * (4 * c + j);
      const float2 z = fbuf[slot(pos_of_freq(f))], zc = fbuf[slot(pos_of_freq((kN - f) & (kN - 1)))];
      A[j] = make_float2((z.x + zc.x) * sa, (z.y - zc.y) * sa * sgn);
      Bv[j] = make_float2((z.y + zc.y) * sa, (zc.x - z.x) * sa * sgn);
    }
    kf_eng[size_t(ha) * (kN / 4) + v] = make_uint4(NT::pack(A[0].x, A[1].x), NT::pack(A[0].y, A[1].y),
                                                    NT::pack(A[2].x, A[3].x), NT::pack(A[2].y, A[3].y));
    if (hb < H)
      kf_eng[size_t(hb) * (kN / 4) + v] = make_uint4(NT::pack(Bv[0].x, Bv[1].x), NT::pack(Bv[0].y, Bv[1].y),
                                                      NT::pack(Bv[2].x, Bv[3].x), NT::pack(Bv[2].y, Bv[3].y));
  }
}

// grid = H.  dk[n] = scale/8192 * Re( c[n] + (fold ? c[fold_off + n] : 0) ),  c = inverse DFT of dk_f, n < Lk.
// Engine order of dk_f (dkf_r128.cuh): index ((qd*128 + k1)*16 + k2l) holds frequency k1 + 128 (16 qd + k2l).
__global__ void __launch_bounds__(kThreads, 3) dk_from_dkf_kernel(const float2* __restrict__ dkf_eng, float* __restrict__ dk, int Lk,
                                                                  float scale, int fold_off, const float2* __restrict__ tw) {
  extern __shared__ float2 fbuf[];
  const int tid = threadIdx.x, h = blockIdx.x;
  const float2* src = dkf_eng + size_t(h) * kN;
  for (int e = tid; e < kN; e += kThreads) {
    const int k2l = e & 15, k1 = (e >> 4) & 127, qd = e >> 11;
    fbuf[slot(k1 + 128 * (16 * qd + k2l))] = src[e];
  }
  __syncthreads();
  fft8192<1>(fbuf, tid, tw);
  const float s = scale * (1.0f / float(kN));
  for (int n = tid; n < Lk; n += kThreads) {
    float c = fbuf[slot(pos_of_freq(n))].x;
    if (fold_off) c += fbuf[slot(pos_of_freq((fold_off + n) & (kN - 1)))].x;
    dk[size_t(h) * Lk + n] = c * s;
  }
}

}  // namespace ffft
}  // namespace bffc

// Outer radix-R stages (R = 2, 4, 8) for N = R x 8192 on CUDA cores: pure streaming, HBM-bound kernels.
//
// Path replaced (reference): the butterfly kernels that wrap the inner Monarch convolution for N > 32K
// (csrc/flashfftconv/butterfly/butterfly_padded_cuda_bf16.cu:17-157 forward, butterfly_padded_ifft_cuda_bf16.cu:
// 15-319 inverse; gated variants :302/:326) — here used already from N = 16K because the fused tcgen05 kernel
// is the 8192-point one.  Same role: outer DFT down the stride-M columns + (R x M) twiddle, with the implicit
// zero padding (rows >= L/M are never read) and the gates applied on load / store.
//
//   n = a*M + n',  k = c + R*k',  M = 8192,  z = u_b + i u_{b+1} (pair packing, see r128_common.cuh)
//   forward : V_c[n'] = W_N^{n' c} * sum_a W_R^{a c} z[a*M + n']          -> planes row ((pair*H + h)*R + c)
//   inverse : z'[a*M + n'] = sum_c W_R^{-a c} W_N^{-n' c} T_c[n']         (1/N is folded into k_f)
// Each thread owns 8 consecutive n' (one 16-byte vector per row).
#pragma once
#include "ptx.cuh"

namespace bffc {
namespace outer {

constexpr int kVec = 8;

template <int kFmt>
DEVINL void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) upk2(Num<kFmt>::unpack(w[i]), f[2 * i], f[2 * i + 1]);
}
template <int kFmt>
DEVINL uint4 pack8(const float (&f)[8]) {
  using NT = Num<kFmt>;
  return make_uint4(NT::pack(f[0], f[1]), NT::pack(f[2], f[3]), NT::pack(f[4], f[5]), NT::pack(f[6], f[7]));
}
template <int kFmt>
DEVINL uint4 hmul8(const uint4& a, const uint4& b) {
  using NT = Num<kFmt>;
  return make_uint4(NT::hmul2(a.x, b.x), NT::hmul2(a.y, b.y), NT::hmul2(a.z, b.z), NT::hmul2(a.w, b.w));
}

struct OuterParams {
  const uint4* u;        // (B, H, L) bf16                                  (real endpoint, level 0)
  const uint4* pregate;  // optional
  const uint4* postgate; // optional
  uint4* y;              // (B, H, L) bf16 (inverse only)
  const uint4* postgate2; // optional second gated output of the inverse: y2 = postgate2 * z' (gated backward: du, dpregate)
  uint4* y2;
  uint4* xre;            // kPlanes: outer-side complex rows (rows x R*M), read by fwd / written by inv
  uint4* xim;
  uint4* pre;            // inner-side planes: real parts,  rows*R rows of M bf16 each
  uint4* pim;            // inner-side planes: imaginary parts
  int B, H, L, pairs;    // this launch covers batch members [0, B) and channels [h0, h0 + H) of tensors with Hs channels
  int Hs, h0;            // channel count of the (B, Hs, L) tensors and first channel of this launch (planes are chunk local)
  int M;                 // inner row length
  float2 step[8];        // exp(-2 pi i t / (R*M)), t = 0..7: neighbour twiddle steps (host computed, double precision)
  int lookahead;         // blocks: a block pulls the input lines of block (its linear id + lookahead) into L2 (0 = off)
  float scale;           // applied to this stage's output (fp16: 1/sqrt(R) per direction; bf16: 1, 1/N lives in k_f)
};

// W_R^{a c} for R <= 8 as exact constants
DEVINL void wr(int R, int e, float& c, float& s) {   // exp(-2 pi i e / R)
  const int t = ((e % R) + R) % R * (8 / R);          // in eighths of a turn
  const float h = 0.70710678118654752f;
  const float cs[8] = {1.f, h, 0.f, -h, -1.f, -h, 0.f, h};
  const float sn[8] = {0.f, -h, -1.f, -h, 0.f, h, 1.f, h};
  c = cs[t]; s = sn[t];
}

DEVINL f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// v += z * exp(-2 pi i e8 / 8)   (e8 = eighths of a turn, a compile-time constant after unrolling)
DEVINL void rot_acc(int e8, f32x2 zr, f32x2 zi, f32x2& vr, f32x2& vi) {
  e8 &= 7;
  const float h = 0.70710678118654752f;
  if (e8 == 0) { vr = add2(vr, zr); vi = add2(vi, zi); }
  else if (e8 == 2) { vr = add2(vr, zi); vi = sub2(vi, zr); }          // * (-i)
  else if (e8 == 4) { vr = sub2(vr, zr); vi = sub2(vi, zi); }          // * (-1)
  else if (e8 == 6) { vr = sub2(vr, zi); vi = add2(vi, zr); }          // * (+i)
  else {
    const float fc = (e8 == 1 || e8 == 7) ? h : -h, fs = (e8 == 1 || e8 == 3) ? -h : h;   // cos, sin of -2 pi e8/8
    const f32x2 c2 = pk2(fc, fc), s2 = pk2(fs, fs), ns2 = pk2(-fs, -fs);
    vr = fma2(zr, c2, fma2(zi, ns2, vr));
    vi = fma2(zr, s2, fma2(zi, c2, vi));
  }
}

template <int kFmt>
DEVINL void unpack8v(const uint4& v, f32x2 (&f)[4]) {
  f[0] = Num<kFmt>::unpack(v.x); f[1] = Num<kFmt>::unpack(v.y); f[2] = Num<kFmt>::unpack(v.z); f[3] = Num<kFmt>::unpack(v.w);
}
template <int kFmt>
DEVINL uint4 pack8v(const f32x2 (&f)[4]) {
  using NT = Num<kFmt>;
  return make_uint4(NT::pack_v(f[0]), NT::pack_v(f[1]), NT::pack_v(f[2]), NT::pack_v(f[3]));
}

// twiddles of 8 consecutive positions: w1[t] = exp(sign * 2 pi i (np + t) / Nl), as 4 packed pairs.
// One accurate sincos for position np, the other seven by the per-level constants step[t] = exp(sign 2 pi i t / Nl).
DEVINL void twiddle8(int np, float inv_nl2, const float2* step, f32x2 (&wc)[4], f32x2 (&ws)[4]) {
  float s0, c0;
  sincospif(float(np) * inv_nl2, &s0, &c0);
  float c[8], s[8];
  c[0] = c0; s[0] = s0;
#pragma unroll
  for (int t = 1; t < 8; ++t) {
    c[t] = c0 * step[t].x - s0 * step[t].y;
    s[t] = c0 * step[t].y + s0 * step[t].x;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) { wc[q] = pk2(c[2 * q], c[2 * q + 1]); ws[q] = pk2(s[2 * q], s[2 * q + 1]); }
}

DEVINL void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Software read-ahead for the streaming level-0 kernels.  A block lives for one load -> compute -> store round trip and
// only 16 warps fit per SM (128 registers), so HBM latency is exposed (ncu: 70 % of the stall samples on the first use
// of the loads, 4.1-4.7 TB/s).  Blocks are dispatched in linear order; each block therefore touches the lines the block
// `lookahead` positions later will load (one thread per 128-byte line), turning those loads into L2 hits.
template <int R, bool kGated>
DEVINL void readahead_level0(const OuterParams& p, bool planes_in) {
  if (p.lookahead <= 0 || (threadIdx.x & 7) != 0) return;
  const unsigned gx = gridDim.x, gy = gridDim.y;
  const unsigned long long total = (unsigned long long)gx * gy * gridDim.z;
  const unsigned long long id = blockIdx.x + (unsigned long long)gx * (blockIdx.y + (unsigned long long)gy * blockIdx.z) + p.lookahead;
  if (id >= total) return;
  const int bx = int(id % gx), h = int((id / gx) % gy), pr = int(id / (gx * (unsigned long long)gy));
  const int np = (bx * blockDim.x + threadIdx.x) * kVec;
  const size_t L8 = size_t(p.L) / kVec;
  const int b0 = 2 * pr, b1 = 2 * pr + 1;
  if (planes_in) {      // inverse: R rows of both planes
#pragma unroll
    for (int c = 0; c < R; ++c) {
      const size_t row = (size_t(pr) * p.H + h) * R + c;
      prefetch_l2(p.pre + row * (p.M / kVec) + np / kVec);
      prefetch_l2(p.pim + row * (p.M / kVec) + np / kVec);
    }
  }
#pragma unroll
  for (int a = 0; a < R; ++a) {
    const int n = a * p.M + np;
    if (n >= p.L) break;
    const size_t o0 = (size_t(b0) * p.Hs + p.h0 + h) * L8 + n / kVec, o1 = (size_t(b1) * p.Hs + p.h0 + h) * L8 + n / kVec;
    if (!planes_in) {
      prefetch_l2(p.u + o0);
      if (b1 < p.B) prefetch_l2(p.u + o1);
    }
    if (kGated) {
      const uint4* g = planes_in ? p.postgate : p.pregate;
      prefetch_l2(g + o0);
      if (b1 < p.B) prefetch_l2(g + o1);
    }
  }
}

// same for the plane-to-plane levels: grid (rows, M / (kVec*blockDim.x), 1), R rows of both planes per block
template <int R>
DEVINL void readahead_planes(const OuterParams& p, bool inverse) {
  if (p.lookahead <= 0 || (threadIdx.x & 7) != 0) return;
  const unsigned gx = gridDim.x;
  const unsigned long long id = blockIdx.x + (unsigned long long)gx * blockIdx.y + p.lookahead;
  if (id >= (unsigned long long)gx * gridDim.y) return;
  const size_t row0 = size_t(id % gx) * R;
  const int np = (int(id / gx) * blockDim.x + threadIdx.x) * kVec;
  const uint4* re = inverse ? p.pre : p.xre;
  const uint4* im = inverse ? p.pim : p.xim;
#pragma unroll
  for (int a = 0; a < R; ++a) {
    const size_t o = ((row0 + a) * p.M + np) / kVec;
    prefetch_l2(re + o);
    prefetch_l2(im + o);
  }
}

// forward: grid (M / (kVec*blockDim.x), H, pairs)   [kPlanes: (rows, M / (kVec*blockDim.x), 1)]
template <int R, bool kGated, bool kPlanes, int kFmt>
__global__ void __launch_bounds__(128, (R <= 4) ? 4 : 2) fwd_kernel(const OuterParams p) {
  const int kM = p.M;
  const int np = ((kPlanes ? blockIdx.y : blockIdx.x) * blockDim.x + threadIdx.x) * kVec;   // n'
  const int h = blockIdx.y, pr = blockIdx.z;
  const int b0 = 2 * pr, b1 = 2 * pr + 1;
  const size_t L8 = size_t(p.L) / kVec;
  if (kPlanes) readahead_planes<R>(p, false); else readahead_level0<R, kGated>(p, false);
  f32x2 zr[R][4], zi[R][4];
  int rows = 0;
#pragma unroll
  for (int a = 0; a < R; ++a) {
    const int n = a * kM + np;
    if (kPlanes) {
      rows = R;
      const size_t o = (size_t(blockIdx.x) * R * kM + n) / kVec;
      unpack8v<kFmt>(__ldg(p.xre + o), zr[a]);
      unpack8v<kFmt>(__ldg(p.xim + o), zi[a]);
    } else if (n < p.L) {
      rows = a + 1;
      const size_t o0 = (size_t(b0) * p.Hs + p.h0 + h) * L8 + n / kVec;
      uint4 v0 = __ldg(p.u + o0);
      if (kGated) v0 = hmul8<kFmt>(v0, __ldg(p.pregate + o0));
      unpack8v<kFmt>(v0, zr[a]);
      if (b1 < p.B) {
        const size_t o1 = (size_t(b1) * p.Hs + p.h0 + h) * L8 + n / kVec;
        uint4 v1 = __ldg(p.u + o1);
        if (kGated) v1 = hmul8<kFmt>(v1, __ldg(p.pregate + o1));
        unpack8v<kFmt>(v1, zi[a]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) zi[a][q] = 0ull;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) { zr[a][q] = 0ull; zi[a][q] = 0ull; }
    }
  }
  f32x2 w1c[4], w1s[4];                      // W_N^{n'+t}
  twiddle8(np, -2.0f / float(R * kM), p.step, w1c, w1s);
  f32x2 wc[4], ws[4];                        // running W_N^{(n'+t) c}, carrying the stage's output scale
#pragma unroll
  for (int q = 0; q < 4; ++q) { wc[q] = pk2(p.scale, p.scale); ws[q] = 0ull; }
#pragma unroll
  for (int c = 0; c < R; ++c) {
    f32x2 vr[4], vi[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { vr[q] = 0ull; vi[q] = 0ull; }
#pragma unroll
    for (int a = 0; a < R; ++a) {
      if (a < rows) {
#pragma unroll
        for (int q = 0; q < 4; ++q) rot_acc((a * c % R) * (8 / R), zr[a][q], zi[a][q], vr[q], vi[q]);
      }
    }
    f32x2 orr[4], oii[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cmul2(vr[q], vi[q], wc[q], ws[q], orr[q], oii[q]);
      if (c + 1 < R) {
        f32x2 nc, ns;
        cmul2(wc[q], ws[q], w1c[q], w1s[q], nc, ns);   // w_{c+1} = w_c * w_1
        wc[q] = nc; ws[q] = ns;
      }
    }
    const size_t row = (kPlanes ? size_t(blockIdx.x) : (size_t(pr) * p.H + h)) * R + c;
    p.pre[row * (kM / kVec) + np / kVec] = pack8v<kFmt>(orr);
    p.pim[row * (kM / kVec) + np / kVec] = pack8v<kFmt>(oii);
  }
}

// inverse: same grid
template <int R, bool kGated, bool kPlanes, int kFmt>
__global__ void __launch_bounds__(128, (R <= 4) ? 4 : 2) inv_kernel(const OuterParams p) {
  const int kM = p.M;
  const int np = ((kPlanes ? blockIdx.y : blockIdx.x) * blockDim.x + threadIdx.x) * kVec;
  const int h = blockIdx.y, pr = blockIdx.z;
  const int b0 = 2 * pr, b1 = 2 * pr + 1;
  const size_t L8 = size_t(p.L) / kVec;
  if (kPlanes) readahead_planes<R>(p, true); else readahead_level0<R, kGated>(p, true);
  f32x2 w1c[4], w1s[4];                      // conj twiddle: exp(+2 pi i (n'+t) / N)
  float2 stepc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) stepc[t] = make_float2(p.step[t].x, -p.step[t].y);
  twiddle8(np, 2.0f / float(R * kM), stepc, w1c, w1s);
  f32x2 wc[4], ws[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { wc[q] = pk2(p.scale, p.scale); ws[q] = 0ull; }
  f32x2 tr[R][4], ti[R][4];
#pragma unroll
  for (int c = 0; c < R; ++c) {
    const size_t row = (kPlanes ? size_t(blockIdx.x) : (size_t(pr) * p.H + h)) * R + c;
    f32x2 xr[4], xi[4];
    unpack8v<kFmt>(__ldg(p.pre + row * (kM / kVec) + np / kVec), xr);
    unpack8v<kFmt>(__ldg(p.pim + row * (kM / kVec) + np / kVec), xi);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cmul2(xr[q], xi[q], wc[q], ws[q], tr[c][q], ti[c][q]);
      if (c + 1 < R) {
        f32x2 nc, ns;
        cmul2(wc[q], ws[q], w1c[q], w1s[q], nc, ns);
        wc[q] = nc; ws[q] = ns;
      }
    }
  }
#pragma unroll
  for (int a = 0; a < R; ++a) {
    const int n = a * kM + np;
    if (kPlanes || n < p.L) {
      f32x2 yr[4], yi[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { yr[q] = 0ull; yi[q] = 0ull; }
#pragma unroll
      for (int c = 0; c < R; ++c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) rot_acc(((R * R - a * c) % R) * (8 / R), tr[c][q], ti[c][q], yr[q], yi[q]);   // W_R^{-ac}
      }
      if (kPlanes) {
        const size_t o = (size_t(blockIdx.x) * R * kM + n) / kVec;
        p.xre[o] = pack8v<kFmt>(yr);
        p.xim[o] = pack8v<kFmt>(yi);
        continue;
      }
      const size_t o0 = (size_t(b0) * p.Hs + p.h0 + h) * L8 + n / kVec;
      uint4 v0 = pack8v<kFmt>(yr);
      if (kGated && p.y2) p.y2[o0] = hmul8<kFmt>(v0, __ldg(p.postgate2 + o0));
      if (kGated) v0 = hmul8<kFmt>(v0, __ldg(p.postgate + o0));
      p.y[o0] = v0;
      if (b1 < p.B) {
        const size_t o1 = (size_t(b1) * p.Hs + p.h0 + h) * L8 + n / kVec;
        uint4 v1 = pack8v<kFmt>(yi);
        if (kGated && p.y2) p.y2[o1] = hmul8<kFmt>(v1, __ldg(p.postgate2 + o1));
        if (kGated) v1 = hmul8<kFmt>(v1, __ldg(p.postgate + o1));
        p.y[o1] = v1;
      }
    }
  }
}

}  // namespace outer
}  // namespace bffc

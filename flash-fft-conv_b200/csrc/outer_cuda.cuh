// Outer radix-R stages (R = 2, 4, 8) for N = R x 8192 on CUDA cores: pure streaming, HBM-bound kernels.
//
// Path replaced (reference): the butterfly kernels that wrap the inner Monarch convolution for N > 32K
// (csrc/flashfftconv/butterfly/butterfly_padded_cuda_bf16.cu:17-157 forward, butterfly_padded_ifft_cuda_bf16.cu:
// 15-319 inverse; gated variants :302/:326) — here used already from N = 16K because the fused tcgen05 kernel
// is the 8192-point one.  Same role: outer DFT down the stride-M columns + (R x M) twiddle, with the implicit
// zero padding (rows >= L/M are never read) and the gates applied on load / store.
//
//   n = a*M + n',  k = c + R*k',  M = 8192,  z = u_b + i u_{b+1} (pair packing, see fwd_r128.cuh)
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
  uint4* xre;            // kPlanes: outer-side complex rows (rows x R*M), read by fwd / written by inv
  uint4* xim;
  uint4* pre;            // inner-side planes: real parts,  rows*R rows of M bf16 each
  uint4* pim;            // inner-side planes: imaginary parts
  int B, H, L, pairs;
  int M;                 // inner row length
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

// forward: grid (M / (kVec*blockDim.x), H, pairs)   [kPlanes: (rows, M / (kVec*blockDim.x), 1)]
template <int R, bool kGated, bool kPlanes, int kFmt>
__global__ void __launch_bounds__(128, (R <= 4) ? 4 : 2) fwd_kernel(const OuterParams p) {
  const int kM = p.M;
  const int np = ((kPlanes ? blockIdx.y : blockIdx.x) * blockDim.x + threadIdx.x) * kVec;   // n'
  const int h = blockIdx.y, pr = blockIdx.z;
  const int b0 = 2 * pr, b1 = 2 * pr + 1;
  const size_t L8 = size_t(p.L) / kVec;
  float zr[R][8], zi[R][8];
  int rows = 0;
#pragma unroll
  for (int a = 0; a < R; ++a) {
    const int n = a * kM + np;
    if (kPlanes) {
      rows = R;
      const size_t o = (size_t(blockIdx.x) * R * kM + n) / kVec;
      unpack8<kFmt>(__ldg(p.xre + o), zr[a]);
      unpack8<kFmt>(__ldg(p.xim + o), zi[a]);
    } else if (n < p.L) {
      rows = a + 1;
      const size_t o0 = (size_t(b0) * p.H + h) * L8 + n / kVec;
      uint4 v0 = __ldg(p.u + o0);
      if (kGated) v0 = hmul8<kFmt>(v0, __ldg(p.pregate + o0));
      unpack8<kFmt>(v0, zr[a]);
      if (b1 < p.B) {
        const size_t o1 = (size_t(b1) * p.H + h) * L8 + n / kVec;
        uint4 v1 = __ldg(p.u + o1);
        if (kGated) v1 = hmul8<kFmt>(v1, __ldg(p.pregate + o1));
        unpack8<kFmt>(v1, zi[a]);
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) zi[a][t] = 0.f;
      }
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) { zr[a][t] = 0.f; zi[a][t] = 0.f; }
    }
  }
  // w1[t] = W_N^{n'+t}
  float w1c[8], w1s[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) sincospif(-2.0f * float(np + t) / float(R * kM), &w1s[t], &w1c[t]);
  float wc[8], ws[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { wc[t] = 1.f; ws[t] = 0.f; }
#pragma unroll
  for (int c = 0; c < R; ++c) {
    float vr[8], vi[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { vr[t] = 0.f; vi[t] = 0.f; }
#pragma unroll
    for (int a = 0; a < R; ++a) {
      if (a < rows) {
        float fc, fs;
        wr(R, a * c, fc, fs);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          vr[t] += zr[a][t] * fc - zi[a][t] * fs;
          vi[t] += zr[a][t] * fs + zi[a][t] * fc;
        }
      }
    }
    float or_[8], oi_[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      or_[t] = (vr[t] * wc[t] - vi[t] * ws[t]) * p.scale;
      oi_[t] = (vr[t] * ws[t] + vi[t] * wc[t]) * p.scale;
      const float nc = wc[t] * w1c[t] - ws[t] * w1s[t];   // w_{c+1} = w_c * w_1
      ws[t] = wc[t] * w1s[t] + ws[t] * w1c[t];
      wc[t] = nc;
    }
    const size_t row = (kPlanes ? size_t(blockIdx.x) : (size_t(pr) * p.H + h)) * R + c;
    p.pre[row * (kM / kVec) + np / kVec] = pack8<kFmt>(or_);
    p.pim[row * (kM / kVec) + np / kVec] = pack8<kFmt>(oi_);
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
  float w1c[8], w1s[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) sincospif(2.0f * float(np + t) / float(R * kM), &w1s[t], &w1c[t]);   // conj twiddle
  float wc[8], ws[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { wc[t] = 1.f; ws[t] = 0.f; }
  float tr[R][8], ti[R][8];
#pragma unroll
  for (int c = 0; c < R; ++c) {
    const size_t row = (kPlanes ? size_t(blockIdx.x) : (size_t(pr) * p.H + h)) * R + c;
    float xr[8], xi[8];
    unpack8<kFmt>(__ldg(p.pre + row * (kM / kVec) + np / kVec), xr);
    unpack8<kFmt>(__ldg(p.pim + row * (kM / kVec) + np / kVec), xi);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      tr[c][t] = xr[t] * wc[t] - xi[t] * ws[t];
      ti[c][t] = xr[t] * ws[t] + xi[t] * wc[t];
      const float nc = wc[t] * w1c[t] - ws[t] * w1s[t];
      ws[t] = wc[t] * w1s[t] + ws[t] * w1c[t];
      wc[t] = nc;
    }
  }
#pragma unroll
  for (int a = 0; a < R; ++a) {
    const int n = a * kM + np;
    if (kPlanes || n < p.L) {
      float yr[8], yi[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) { yr[t] = 0.f; yi[t] = 0.f; }
#pragma unroll
      for (int c = 0; c < R; ++c) {
        float fc, fs;
        wr(R, -a * c, fc, fs);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          yr[t] += tr[c][t] * fc - ti[c][t] * fs;
          yi[t] += tr[c][t] * fs + ti[c][t] * fc;
        }
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) { yr[t] *= p.scale; yi[t] *= p.scale; }
      if (kPlanes) {
        const size_t o = (size_t(blockIdx.x) * R * kM + n) / kVec;
        p.xre[o] = pack8<kFmt>(yr);
        p.xim[o] = pack8<kFmt>(yi);
        continue;
      }
      const size_t o0 = (size_t(b0) * p.H + h) * L8 + n / kVec;
      uint4 v0 = pack8<kFmt>(yr);
      if (kGated) v0 = hmul8<kFmt>(v0, __ldg(p.postgate + o0));
      p.y[o0] = v0;
      if (b1 < p.B) {
        const size_t o1 = (size_t(b1) * p.H + h) * L8 + n / kVec;
        uint4 v1 = pack8<kFmt>(yi);
        if (kGated) v1 = hmul8<kFmt>(v1, __ldg(p.postgate + o1));
        p.y[o1] = v1;
      }
    }
  }
}

}  // namespace outer
}  // namespace bffc

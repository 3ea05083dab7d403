// Backward filter-gradient kernel, N = 128 x 64 (= 8192), bf16, sm_100a.
//
// Path replaced (reference): the dk_f part of monarch_conv_bwd_cuda_kernel
// (kernels_bf16/monarch_cuda_32_16_16_bwd_kernel_bf16.h:505-509,571-581,711-734: D = FFT(dout), X = FFT(u),
// dk_f partial = sum over the CTA's batch tile of D * conj(X)) plus the host-side `dk_f_out.sum(0)` over the
// (B/Bt, H, N, 2) bf16 partials (monarch_cuda_interface_bwd_bf16.cu:820,1107).  Here the sum over the whole
// batch is accumulated in fp32 registers inside one CTA per channel; nothing but the final (H, N) complex
// fp32 gradient is written.
//
// Pair packing: z_u = u_b + i u_{b+1}, z_d = dout_b + i dout_{b+1}.  FFT(z_d) * conj(FFT(z_u)) is the spectrum
// of corr(d_b,u_b) + corr(d_{b+1},u_{b+1}) + i (cross terms); the cross terms are purely imaginary in the time
// domain, and the caller takes the real part of the inverse FFT (as the reference does, conv.py:1817-1820),
// so summing the packed products over pairs gives exactly dk.  An odd batch is completed with an all-zero
// partner (TMA out-of-bounds fill).
//
// Machine mapping: pipeline 0 (threads 0..255) transforms the u pair, pipeline 1 the dout pair, with the
// same stage-1 / pass-1 / stage-2 chain as the forward kernel (fwd_r128.cuh); then all 512 threads read
// both spectra from TMEM and accumulate 16 complex products each.
#pragma once
#include "fwd_r128.cuh"

namespace bffc {

struct DkfParams {
  const __nv_bfloat16* dftC;
  const __nv_bfloat16* dftS;
  const uint8_t* gtiles;
  float2* dkf;               // [H][4][128][16] complex fp32: k2 = 16*q + t, frequency k = k1 + 128*k2
  int B, H, L, pairs, kmask;  // pairs = batch groups per channel; kmask as in FwdParams
  int nseg, seg_bytes;        // segmented tiles (small sizes), see load_tile()
  float tw_scale;            // see FwdParams::tw_scale; dkf_unpack compensates
  int gated;                 // 1: u is multiplied by pregate and dout by postgate on load (tm_ui / tm_di = gate maps)
};

namespace r128 {

constexpr int kSmemTotalDkf = kSmemData + kSmemG + kSmemBars + 1024;
constexpr int kSmemTotalDkfGated = kSmemTotalDkf + kSmemGate + 1024;

// kPlanes: inputs are complex rows in bf16 planes (composite sizes): (tm_u, tm_ui) = real / imaginary plane of the
// transformed u rows, (tm_d, tm_di) likewise for dout; p.H = number of k_f rows, row = pr * p.H + channel.
template <bool kPlanes, int kFmt = 1>
__global__ void __launch_bounds__(kThreads, 1)
dkf_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_d,
           const __grid_constant__ CUtensorMap tm_ui, const __grid_constant__ CUtensorMap tm_di, const DkfParams p) {
  using NT = Num<kFmt>;
  constexpr uint32_t ID_N128_MN = Idesc<kFmt>::N128_MN, ID_N64_MN = Idesc<kFmt>::N64_MN, ID_N64_MN_NEG = Idesc<kFmt>::N64_MN_NEG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_g = sbase + kSmemData;
  const uint32_t s_bars = s_g + kSmemG;
  const uint32_t s_gate0 = s_bars + kSmemBars + 960;   // gated only (1024-byte aligned)
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));

  const int tid = threadIdx.x;
  const int pipe = __shfl_sync(0xffffffffu, tid >> 8, 0);   // 0: u, 1: dout; warp-uniform for the compiler (uniform-register MMA issue)
  const int half = (tid >> 7) & 1;
  const int lane = tid & 127;
  const int warp_q = (tid >> 5) & 3;
  const bool lead_warp = ((tid & 255) < 32);

  const uint32_t bar_tma0 = s_bars + pipe * 24;
  const uint32_t bar_mma = s_bars + pipe * 24 + 16;
  const uint32_t s_tmemptr = s_bars + 48;

  if (tid == 0) {
    tma_prefetch_desc(&tm_u);
    tma_prefetch_desc(&tm_d);
    if (kPlanes) { tma_prefetch_desc(&tm_ui); tma_prefetch_desc(&tm_di); }
  }
  if ((tid & 255) == 0) {
    mbar_init(bar_tma0, 1);
    mbar_init(bar_tma0 + 8, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (tid < 32) {
    tmem_alloc(s_tmemptr, 512);
    tmem_relinquish();
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.gtiles);
    uint4* dst = reinterpret_cast<uint4*>(gen_base + kSmemData);
    for (int i = tid; i < kSmemG / 16; i += kThreads) dst[i] = src[i];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + kSmemData + kSmemG + 48);
  const uint32_t tlane = tmem_base + (uint32_t(warp_q * 32) << 16);
  {
    const uint4* row = reinterpret_cast<const uint4*>((pipe == 0 ? p.dftC : p.dftS) + lane * 128) + half * 8;
    const uint32_t tcol = tlane + (pipe == 0 ? kColC : kColS) + 32 * half;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t v[16];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        uint4 w = row[q * 4 + r];
        v[4 * r + 0] = w.x; v[4 * r + 1] = w.y; v[4 * r + 2] = w.z; v[4 * r + 3] = w.w;
      }
      tmem_st16(tcol + 16 * q, v);
    }
    tmem_st_wait();
  }
  __half2 twc[16], tws[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    float s0, c0, s1, c1;
    sincospif(-2.0f * float((lane * (32 * half + 2 * q)) & 8191) / 8192.0f, &s0, &c0);
    sincospif(-2.0f * float((lane * (32 * half + 2 * q + 1)) & 8191) / 8192.0f, &s1, &c1);
    twc[q] = __floats2half2_rn(c0 * p.tw_scale, c1 * p.tw_scale);
    tws[q] = __floats2half2_rn(s0 * p.tw_scale, s1 * p.tw_scale);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t s_slot0 = sbase + pipe * 2 * kSlotBytes;
  const uint32_t tD = tlane + colD(pipe);
  const uint32_t tA = tlane + colA(pipe);
  const uint32_t tDu = tlane + colD(0), tDd = tlane + colD(1);
  const uint32_t tD0 = tmem_base + colD(pipe);
  const uint32_t tA0 = tmem_base + colA(pipe);
  const uint32_t tC0 = tmem_base + kColC;
  const uint32_t tS0 = tmem_base + kColS;
  const uint32_t bar_id = 1 + pipe;
  const uint32_t sG0 = s_g;
  const CUtensorMap* tm = (pipe == 0) ? &tm_u : &tm_d;
  const CUtensorMap* tmi = (kPlanes || p.gated) ? ((pipe == 0) ? &tm_ui : &tm_di) : tm;   // imaginary plane / gate

  // units of this CTA: (h, pr) for h = blockIdx.x, blockIdx.x + gridDim.x, ...
  const int nh = (p.H - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int n_units = nh * p.pairs;
  auto unit_h = [&](int n) { return int(blockIdx.x) + (n / p.pairs) * int(gridDim.x); };
  auto issue_load = [&](int n, int slot) {
    const int h = unit_h(n), pr = n % p.pairs;
    const uint32_t bar = bar_tma0 + 8 * slot;
    const uint32_t dst = s_slot0 + slot * kSlotBytes;
    mbar_expect_tx(bar, (!kPlanes && p.gated) ? 2 * kSlotBytes : kSlotBytes);
    if (kPlanes) {
      tma_load_3d(dst, tm, bar, 0, 0, pr * p.H + h);
      tma_load_3d(dst + kTileBytes, tmi, bar, 0, 0, pr * p.H + h);
    } else {
      load_tile(dst, tm, bar, p.B, p.H, h, pr, 0, p.nseg, p.seg_bytes);              // members beyond the batch: zeros
      load_tile(dst + kTileBytes, tm, bar, p.B, p.H, h, pr, 1, p.nseg, p.seg_bytes);
      if (p.gated) {       // gate tiles (pregate for u, postgate for dout) into this pipeline's gate slot
        const uint32_t gd = s_gate0 + pipe * kSlotBytes;
        load_tile(gd, tmi, bar, p.B, p.H, h, pr, 0, p.nseg, p.seg_bytes);
        load_tile(gd + kTileBytes, tmi, bar, p.B, p.H, h, pr, 1, p.nseg, p.seg_bytes);
      }
    }
  };
  uint32_t mma_phase = 0;
  auto wait_mma = [&]() {
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  };

  if (lead_warp && n_units > 0) {
    if (elect_one()) issue_load(0, 0);
    __syncwarp();
  }

  f32x2 acc_r[8], acc_i[8];      // 16 complex accumulators: k2 = 16*(2*pipe+half) + 2*q + {0,1}
#pragma unroll
  for (int q = 0; q < 8; ++q) { acc_r[q] = 0ull; acc_i[q] = 0ull; }

  for (int n = 0; n < n_units; ++n) {
    const int slot = n & 1;
    const uint32_t sX = s_slot0 + slot * kSlotBytes;
    if (!kPlanes && p.gated) {
      // pass 0: X <- bf16(x * gate) in shared memory (reference: gated loads of the bwd kernel,
      // kernels_bf16/monarch_cuda_32_16_16_bwd_kernel_bf16.h:505-509,571-581)
      mbar_wait(bar_tma0 + 8 * slot, (n >> 1) & 1);
      const uint32_t sG = s_gate0 + pipe * kSlotBytes;
#pragma unroll
      for (int part = 0; part < 2; ++part)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t off = part * kTileBytes + uint32_t(lane) * 128u + uint32_t(4 * half + c) * 16u;
          const uint4 a = ld_shared_v4(sX + off), g = ld_shared_v4(sG + off);
          st_shared_v4(sX + off, NT::hmul2(a.x, g.x), NT::hmul2(a.y, g.y), NT::hmul2(a.z, g.z), NT::hmul2(a.w, g.w));
        }
      fence_proxy_async_smem();
      named_bar_sync(bar_id, kPipeThreads);
    }
    // ---------------- stage 1
    if (lead_warp) {
      if (!(!kPlanes && p.gated)) mbar_wait(bar_tma0 + 8 * slot, (n >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dXr = tile_desc(sX), dXi = tile_desc(sX + kTileBytes);   // a K step = +(2048 >> 4) in the address field
        if (p.kmask == 0xff) {
#pragma unroll
          for (int s = 0; s < 8; ++s) mma_ts(tD0, tC0 + 8 * s, dXr + 128 * s, ID_N128_MN, s > 0);
#pragma unroll
          for (int s = 0; s < 8; ++s) mma_ts(tD0, tS0 + 8 * s, dXi + 128 * s, ID_N64_MN, 1);
#pragma unroll
          for (int s = 0; s < 8; ++s) mma_ts(tD0 + 64, tS0 + 8 * s, dXr + 128 * s, ID_N64_MN_NEG, 1);
        } else {
          uint32_t acc = 0;
          for (int s = 0; s < 8; ++s)
            if ((p.kmask >> s) & 1) { mma_ts(tD0, tC0 + 8 * s, dXr + 128 * s, ID_N128_MN, acc); acc = 1; }
          for (int s = 0; s < 8; ++s)
            if ((p.kmask >> s) & 1) mma_ts(tD0, tS0 + 8 * s, dXi + 128 * s, ID_N64_MN, 1);
          for (int s = 0; s < 8; ++s)
            if ((p.kmask >> s) & 1) mma_ts(tD0 + 64, tS0 + 8 * s, dXr + 128 * s, ID_N64_MN_NEG, 1);
        }
        mma_commit(bar_mma);
        if (n + 1 < n_units) issue_load(n + 1, slot ^ 1);   // other slot: its stage 1 finished a unit ago
      }
      __syncwarp();
    }
    wait_mma();
    // ---------------- pass 1
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      uint32_t re[16], im[16];
      tmem_ld16(tD + 32 * half + 16 * sub, re);
      tmem_ld16(tD + 64 + 32 * half + 16 * sub, im);
      tmem_ld_wait();
      reg_fence(re); reg_fence(im);
      uint32_t ore[8], oim[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float2 c = __half22float2(twc[8 * sub + q]), sn = __half22float2(tws[8 * sub + q]);
        f32x2 vr, vi;
        cmul2(pk2u(re[2 * q], re[2 * q + 1]), pk2u(im[2 * q], im[2 * q + 1]), pk2(c.x, c.y), pk2(sn.x, sn.y), vr, vi);
        ore[q] = NT::pack_v(vr);
        oim[q] = NT::pack_v(vi);
      }
      tmem_st8(tA + 16 * half + 8 * sub, ore);
      tmem_st8(tA + 32 + 16 * half + 8 * sub, oim);
    }
    tmem_st_wait();
    tc_fence_before();
    named_bar_sync(bar_id, kPipeThreads);
    // ---------------- stage 2
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dG0 = pair_desc(sG0, 8192), dG1 = pair_desc(sG0 + 16384, 8192);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ts(tD0, tA0 + 8 * s, dG0 + 128 * s, ID_N128_MN, s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ts(tD0, tA0 + 32 + 8 * s, dG1 + 128 * s, ID_N128_MN, 1);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    wait_mma();
    // ---------------- both spectra are in TMEM: accumulate Zd * conj(Zu) over this thread's 16 frequencies
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    {
      const int qd = 2 * pipe + half;
      uint32_t ur[16], ui[16], dr[16], di[16];
      tmem_ld16(tDu + 16 * qd, ur);
      tmem_ld16(tDu + 64 + 16 * qd, ui);
      tmem_ld16(tDd + 16 * qd, dr);
      tmem_ld16(tDd + 64 + 16 * qd, di);
      tmem_ld_wait();
      reg_fence(ur); reg_fence(ui); reg_fence(dr); reg_fence(di);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const f32x2 a = pk2u(dr[2 * q], dr[2 * q + 1]), b = pk2u(di[2 * q], di[2 * q + 1]);
        const f32x2 c = pk2u(ur[2 * q], ur[2 * q + 1]), d = pk2u(ui[2 * q], ui[2 * q + 1]);
        // (a + ib)(c - id) = (ac + bd) + i(bc - ad)
        acc_r[q] = fma2(a, c, fma2(b, d, acc_r[q]));
        acc_i[q] = fma2(b, c, acc_i[q]);
        acc_i[q] = sub2(acc_i[q], mul2(a, d));
      }
    }
    tc_fence_before();
    __syncthreads();       // D regions may be overwritten by the next stage 1
    tc_fence_after();
    // ---------------- channel finished: write its gradient spectrum
    if ((n + 1) % p.pairs == 0) {
      const int h = unit_h(n);
      const int qd = 2 * pipe + half;
      float4* out = reinterpret_cast<float4*>(p.dkf + ((size_t(h) * 4 + qd) * 128 + lane) * 16);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float r0, r1, i0, i1;
        upk2(acc_r[q], r0, r1);
        upk2(acc_i[q], i0, i1);
        out[q] = make_float4(r0, i0, r1, i1);
        acc_r[q] = 0ull; acc_i[q] = 0ull;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

// dk_f engine order -> natural order complex64 (reference analogue: the inverse permutation at conv.py:1818).
// Composite sizes: channel row (h*R0 + c0)*R1 + c1 holds frequencies k = c0 + R0*(c1 + R1*(k1 + 128*k2)).
// One thread moves the 16 consecutive-k2 values of one (row, quarter, k1): 128 contiguous bytes in, 16 stores that are
// contiguous across the k1 lanes of a warp.
__global__ void dkf_unpack_kernel(const float2* __restrict__ eng, float2* __restrict__ nat, int N, int R0, int R1,
                                  float scale) {
  const int h = blockIdx.y;
  const int R = R0 * R1;
  const int ngroups = R * 4 * 128;                      // (row, quarter, k1) groups per channel
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
    const int k1 = g & 127, qd = (g >> 7) & 3, row = g >> 9;
    const int c0 = row / R1, c1 = row % R1;
    const float4* in = reinterpret_cast<const float4*>(eng + ((size_t(h) * R + row) * 4 + qd) * 128 * 16 + size_t(k1) * 16);
#pragma unroll
    for (int t2 = 0; t2 < 8; ++t2) {
      const float4 v = in[t2];
      const int k2 = 16 * qd + 2 * t2;
      const size_t ka = size_t(c0) + size_t(R0) * (c1 + size_t(R1) * (k1 + 128 * k2));
      const size_t kb = size_t(c0) + size_t(R0) * (c1 + size_t(R1) * (k1 + 128 * (k2 + 1)));
      nat[size_t(h) * N + ka] = make_float2(v.x * scale, v.y * scale);
      nat[size_t(h) * N + kb] = make_float2(v.z * scale, v.w * scale);
    }
  }
}

}  // namespace r128
}  // namespace bffc

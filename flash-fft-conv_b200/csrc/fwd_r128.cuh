// Fused forward FFT-convolution kernel, N = 128 x 64 (= 8192), bf16, sm_100a.
//
// Path replaced (reference): monarch_conv_cuda_kernel<32,8,8192,...>
// (csrc/flashfftconv/monarch_cuda/kernels_bf16/monarch_cuda_32_16_16_kernel_bf16.h:15-801) and its
// launcher (monarch_cuda_interface_fwd_bf16.cu:656-760).  Same math, different machine mapping:
//
//  * two real sequences (b, b+1) of one channel h are packed as ONE complex sequence z = u_b + i u_{b+1};
//    conv(z, k) = conv(u_b,k) + i conv(u_{b+1},k) because k is real, so no Hermitian split is needed.
//  * N = 128 * 64, n = i*64 + j.  Stage 1 contracts i with the 128x128 DFT matrix as the tcgen05 A operand
//    (resident in TMEM for the whole kernel) and the TMA-loaded (128 x 64) input tile as an MN-major B
//    operand: D1[k1, j] lands in TMEM with lane = k1.
//  * everything between stage 1 and the last stage is "row local": lane k1 owns the 64-point transform
//    over j, done as 8 x 8 with tiny (N=16,K=16) MMAs whose A operand is re-packed in TMEM by the owning
//    thread (tcgen05.ld -> twiddle in fp32 registers -> bf16x2 -> tcgen05.st).  The inner twiddles
//    W_64^{a*j2} are folded into the per-block B matrices; the lane dependent twiddle W_N^{k1*j} is
//    factored as W^{8*k1*j1} (pass 1) * W^{k1*j2} (pass 2) so each thread keeps only 16 complex factors.
//  * the last stage contracts k1 (the lane index), so its input is written to shared memory as the
//    MN-major B operand; the output accumulator has lane = i and is stored with TMA.
//  * a CTA runs two independent "pipelines" (warpgroups); while one waits on its MMAs the other runs its
//    CUDA-core pass.  No intermediate ever touches HBM.
#pragma once
#include "ptx.cuh"
#include <cuda.h>

namespace bffc {

struct FwdParams {
  const uint32_t* kf;        // [H][128][64] packed (re | im<<16) bf16, engine order, scaled 1/N
  const __nv_bfloat16* dftC; // [128][128] cos(2*pi*m*k/128)
  const __nv_bfloat16* dftS; // [128][128] sin(2*pi*m*k/128)
  const uint8_t* bsmall;     // kNumSmall x 512 B, canonical no-swizzle K-major 16x16 bf16 B matrices
  int B, H;                  // batch, channels
  int pairs;                 // ceil(B/2)
  int ksteps;                // number of 16-row K steps of the input tile that are non-zero (L/64/16 up)
  int units;                 // H * pairs
  float* dbg;                // optional stage dump [stage][128][128]
  int dbg_stages;
};

namespace r128 {

constexpr int kThreads = 256;
constexpr int kTileBytes = 128 * 128;          // one (128 rows x 64 bf16) tile
constexpr int kSlotBytes = 2 * kTileBytes;     // re tile + im tile
constexpr int kNumSmall = 18;                  // B2a, B2b[8], B3b[8], B3a
constexpr int kSmallBytes = 512;
constexpr int kSmemData = 4 * kSlotBytes;      // 2 pipelines x 2 slots
constexpr int kSmemSmall = kNumSmall * kSmallBytes;
constexpr int kSmemBars = 64;
constexpr int kSmemTotal = kSmemData + kSmemSmall + kSmemBars + 1024;  // + alignment slack

// TMEM columns
constexpr uint32_t kColC = 0, kColS = 64;                 // DFT cos / sin, bf16 K-major A operand
DEVINL constexpr uint32_t colD(int pipe) { return 128 + 192 * pipe; }        // 128 fp32 cols
DEVINL constexpr uint32_t colA(int pipe) { return 128 + 192 * pipe + 128; }  // 64 cols (bf16x2)

constexpr uint32_t ID_N128_MN = make_idesc(1, 128, true, false);
constexpr uint32_t ID_N64_MN = make_idesc(1, 64, true, false);
constexpr uint32_t ID_N64_MN_NEG = make_idesc(1, 64, true, true);
constexpr uint32_t ID_N16_K = make_idesc(1, 16, false, false);

DEVINL void cmul(float ar, float ai, float br, float bi, float& cr, float& ci) {
  cr = ar * br - ai * bi;
  ci = ar * bi + ai * br;
}

template <bool kDebug>
__global__ void __launch_bounds__(kThreads, 1)
fwd_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_y, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_small = sbase + kSmemData;
  const uint32_t s_bars = s_small + kSmemSmall;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));  // generic pointer to aligned base

  const int tid = threadIdx.x;
  const int pipe = tid >> 7;           // warpgroup = pipeline
  const int lane = tid & 127;          // TMEM lane owned by this thread (= k1, later = i)
  const int warp_q = (tid >> 5) & 3;   // TMEM sub-partition of this warp
  const bool leader = (lane == 0);

  const uint32_t bar_tma0 = s_bars + pipe * 24;       // two TMA barriers
  const uint32_t bar_mma = s_bars + pipe * 24 + 16;   // one MMA barrier
  const uint32_t s_tmemptr = s_bars + 48;

  // ---------------------------------------------------------------- prologue
  if (tid == 0) {
    tma_prefetch_desc(&tm_u);
    tma_prefetch_desc(&tm_y);
  }
  if (leader) {
    mbar_init(bar_tma0, 1);
    mbar_init(bar_tma0 + 8, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (tid < 32) {
    tmem_alloc(s_tmemptr, 512);
    tmem_relinquish();
  }
  // small B matrices -> smem (generic proxy writes, later read by the MMA/async proxy)
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.bsmall);
    uint4* dst = reinterpret_cast<uint4*>(gen_base + kSmemData);
    for (int i = tid; i < kSmemSmall / 16; i += kThreads) dst[i] = src[i];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + kSmemData + kSmemSmall + 48);
  const uint32_t tlane = tmem_base + (uint32_t(warp_q * 32) << 16);  // this warp's lane window

  // DFT matrices -> TMEM (pipeline 0 loads cos, pipeline 1 loads sin); row = lane, 128 bf16 = 64 cols
  {
    const uint4* row = reinterpret_cast<const uint4*>((pipe == 0 ? p.dftC : p.dftS) + lane * 128);
    const uint32_t tcol = tlane + (pipe == 0 ? kColC : kColS);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
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

  // lane dependent twiddles, factored:  W_N^{k1*j} = twA[j1] * twB[j2],  j = 8*j1 + j2
  float twAr[8], twAi[8], twBr[8], twBi[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float s, c;
    sincospif(-2.0f * float((lane * 8 * t) & 8191) / 8192.0f, &s, &c);
    twAr[t] = c; twAi[t] = s;
    sincospif(-2.0f * float(lane * t) / 8192.0f, &s, &c);
    twBr[t] = c; twBi[t] = s;
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---------------------------------------------------------------- work assignment
  const int gp = blockIdx.x * 2 + pipe;
  const int GP = gridDim.x * 2;
  const int u_begin = int((long long)p.units * gp / GP);
  const int u_end = int((long long)p.units * (gp + 1) / GP);

  const uint32_t s_slot0 = sbase + pipe * 2 * kSlotBytes;
  const uint32_t tD = tlane + colD(pipe);
  const uint32_t tA = tlane + colA(pipe);
  const uint32_t tD0 = tmem_base + colD(pipe);   // lane 0 addresses for the MMA issuer
  const uint32_t tA0 = tmem_base + colA(pipe);
  const uint32_t tC0 = tmem_base + kColC;
  const uint32_t tS0 = tmem_base + kColS;
  const uint32_t bar_id = 1 + pipe;

  auto seq_index = [&](int unit, int which) {   // global sequence index (b*H + h) of the re / im member
    const int h = unit / p.pairs, pr = unit - h * p.pairs;
    int b = 2 * pr + which;
    if (b >= p.B) b = p.B - 1;                  // odd batch: duplicate, result discarded
    return b * p.H + h;
  };
  auto issue_load = [&](int unit, int slot) {
    const uint32_t bar = bar_tma0 + 8 * slot;
    const uint32_t dst = s_slot0 + slot * kSlotBytes;
    mbar_expect_tx(bar, kSlotBytes);
    tma_load_3d(dst, &tm_u, bar, 0, 0, seq_index(unit, 0));
    tma_load_3d(dst + kTileBytes, &tm_u, bar, 0, 0, seq_index(unit, 1));
  };

  uint32_t mma_phase = 0;
  auto wait_mma = [&]() {
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  };
  // all 128 threads finished writing TMEM (st) / smem; hand over to the MMA issuer
  auto sync_pipe_tmem = [&]() {
    tmem_st_wait();
    tc_fence_before();
    named_bar_sync(bar_id, 128);
  };
  int dbg_stage = 0;
  auto dump = [&](bool first) {
    if (kDebug) {
      if (first && p.dbg != nullptr && dbg_stage < p.dbg_stages) {
        float* o = p.dbg + (size_t(dbg_stage) * 128 + lane) * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(tD + 32 * c, v);
          tmem_ld_wait();
          reg_fence(v);
#pragma unroll
          for (int t = 0; t < 32; ++t) o[32 * c + t] = __uint_as_float(v[t]);
        }
      }
      ++dbg_stage;
    }
  };

  if (leader && u_begin < u_end) issue_load(u_begin, 0);

  for (int unit = u_begin, n = 0; unit < u_end; ++unit, ++n) {
    const int slot = n & 1;
    const uint32_t sX = s_slot0 + slot * kSlotBytes;   // re tile; im tile at +kTileBytes
    const bool first = kDebug && (unit == 0);
    const int h = unit / p.pairs;

    // ---------------- stage 1: D1 = F128 * X   (lane = k1, cols [0,64) re, [64,128) im)
    if (leader) {
      mbar_wait(bar_tma0 + 8 * slot, (n >> 1) & 1);
      tc_fence_after();
      // D[:,0:128]  = C * [Xr | Xi]
      for (int s = 0; s < p.ksteps; ++s)
        mma_ts(tD0, tC0 + 8 * s, make_sdesc(sX + s * 2048, kTileBytes, 1024, 2), ID_N128_MN, s > 0);
      // D[:,0:64]  += S * Xi ;  D[:,64:128] += (-S) * Xr        (F = C - iS)
      for (int s = 0; s < p.ksteps; ++s)
        mma_ts(tD0, tS0 + 8 * s, make_sdesc(sX + kTileBytes + s * 2048, kTileBytes, 1024, 2), ID_N64_MN, 1);
      for (int s = 0; s < p.ksteps; ++s)
        mma_ts(tD0 + 64, tS0 + 8 * s, make_sdesc(sX + s * 2048, kTileBytes, 1024, 2), ID_N64_MN_NEG, 1);
      mma_commit(bar_mma);
      if (unit + 1 < u_end) {      // prefetch next unit into the other slot (its last reader: TMA store)
        tma_store_wait_read0();
        issue_load(unit + 1, slot ^ 1);
      }
    }
    wait_mma();
    dump(first);

    // ---------------- pass 1: * W^{8*k1*j1}, pack A1: block j2 = cols [8*j2, 8*j2+8)
    //   K order inside block: c-major, [re j1=4c..4c+3 | im j1=4c..4c+3]
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t re[32], im[32];
      tmem_ld32(tD + 32 * c, re);
      tmem_ld32(tD + 64 + 32 * c, im);
      tmem_ld_wait();
      reg_fence(re); reg_fence(im);
#pragma unroll
      for (int j2 = 0; j2 < 8; ++j2) {
        float vr[4], vi[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          cmul(__uint_as_float(re[8 * r + j2]), __uint_as_float(im[8 * r + j2]), twAr[4 * c + r], twAi[4 * c + r],
               vr[r], vi[r]);
        tmem_st4(tA + 8 * j2 + 4 * c, pack_bf16x2(vr[0], vr[1]), pack_bf16x2(vr[2], vr[3]),
                 pack_bf16x2(vi[0], vi[1]), pack_bf16x2(vi[2], vi[3]));
      }
    }
    sync_pipe_tmem();
    // ---------------- stage 2a: contract j1 (8 blocks, one per j2)  D block j2 = cols [16*j2,+16) [re a | im a]
    if (leader) {
      tc_fence_after();
      const uint64_t b2a = make_sdesc(s_small + 0 * kSmallBytes, 128, 256, 0);
#pragma unroll
      for (int j2 = 0; j2 < 8; ++j2) mma_ts(tD0 + 16 * j2, tA0 + 8 * j2, b2a, ID_N16_K, 0);
      mma_commit(bar_mma);
    }
    wait_mma();
    dump(first);

    // ---------------- pass 2: * W^{k1*j2}; regroup to blocks by a: A2 block a = [re j2 0..7 | im j2 0..7]
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t re[8][4], im[8][4];
#pragma unroll
      for (int j2 = 0; j2 < 8; ++j2) {
        tmem_ld4(tD + 16 * j2 + 4 * c, re[j2]);
        tmem_ld4(tD + 16 * j2 + 8 + 4 * c, im[j2]);
      }
      tmem_ld_wait();
      reg_fence(re); reg_fence(im);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float vr[8], vi[8];
#pragma unroll
        for (int j2 = 0; j2 < 8; ++j2)
          cmul(__uint_as_float(re[j2][a]), __uint_as_float(im[j2][a]), twBr[j2], twBi[j2], vr[j2], vi[j2]);
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          o[q] = pack_bf16x2(vr[2 * q], vr[2 * q + 1]);
          o[4 + q] = pack_bf16x2(vi[2 * q], vi[2 * q + 1]);
        }
        tmem_st8(tA + 8 * (4 * c + a), o);
      }
    }
    sync_pipe_tmem();
    // ---------------- stage 2b: contract j2 (block a uses its own B with W_64^{a*j2} folded in)
    if (leader) {
      tc_fence_after();
#pragma unroll
      for (int a = 0; a < 8; ++a)
        mma_ts(tD0 + 16 * a, tA0 + 8 * a, make_sdesc(s_small + (1 + a) * kSmallBytes, 128, 256, 0), ID_N16_K, 0);
      mma_commit(bar_mma);
    }
    // k_f row of this lane: 64 packed complex; issue the loads before blocking on the MMA barrier
    const uint4* kfrow = reinterpret_cast<const uint4*>(p.kf + (size_t(h) * 128 + lane) * 64);
    wait_mma();
    dump(first);

    // ---------------- pass 3: * k_f  (frequency k = k1 + 128*(a + 8*d)); same block layout in and out
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      uint32_t d[16];
      tmem_ld16(tD + 16 * a, d);
      const uint4 k0 = __ldg(kfrow + 2 * a), k1v = __ldg(kfrow + 2 * a + 1);
      const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1v.x, k1v.y, k1v.z, k1v.w};
      tmem_ld_wait();
      reg_fence(d);
      float vr[8], vi[8];
#pragma unroll
      for (int t = 0; t < 8; ++t)
        cmul(__uint_as_float(d[t]), __uint_as_float(d[8 + t]), __uint_as_float(kw[t] << 16),
             __uint_as_float(kw[t] & 0xffff0000u), vr[t], vi[t]);
      uint32_t o[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        o[q] = pack_bf16x2(vr[2 * q], vr[2 * q + 1]);
        o[4 + q] = pack_bf16x2(vi[2 * q], vi[2 * q + 1]);
      }
      tmem_st8(tA + 8 * a, o);
    }
    sync_pipe_tmem();
    // ---------------- stage 3b: inverse of 2b (contract d -> j2)
    if (leader) {
      tc_fence_after();
#pragma unroll
      for (int a = 0; a < 8; ++a)
        mma_ts(tD0 + 16 * a, tA0 + 8 * a, make_sdesc(s_small + (9 + a) * kSmallBytes, 128, 256, 0), ID_N16_K, 0);
      mma_commit(bar_mma);
    }
    wait_mma();
    dump(first);

    // ---------------- pass 4: * conj W^{k1*j2}; regroup to blocks by j2: A4 block j2 = [re a | im a]
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t re[8][4], im[8][4];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        tmem_ld4(tD + 16 * a + 4 * c, re[a]);
        tmem_ld4(tD + 16 * a + 8 + 4 * c, im[a]);
      }
      tmem_ld_wait();
      reg_fence(re); reg_fence(im);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j2 = 4 * c + t;
        float vr[8], vi[8];
#pragma unroll
        for (int a = 0; a < 8; ++a)
          cmul(__uint_as_float(re[a][t]), __uint_as_float(im[a][t]), twBr[j2], -twBi[j2], vr[a], vi[a]);
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          o[q] = pack_bf16x2(vr[2 * q], vr[2 * q + 1]);
          o[4 + q] = pack_bf16x2(vi[2 * q], vi[2 * q + 1]);
        }
        tmem_st8(tA + 8 * j2, o);
      }
    }
    sync_pipe_tmem();
    // ---------------- stage 3a: inverse of 2a (contract a -> j1), D block j2 = [re j1 | im j1]
    if (leader) {
      tc_fence_after();
      const uint64_t b3a = make_sdesc(s_small + 17 * kSmallBytes, 128, 256, 0);
#pragma unroll
      for (int j2 = 0; j2 < 8; ++j2) mma_ts(tD0 + 16 * j2, tA0 + 8 * j2, b3a, ID_N16_K, 0);
      mma_commit(bar_mma);
    }
    wait_mma();
    dump(first);

    // ---------------- pass 5: * conj W^{8*k1*j1}; write rows k1 of the B operand [Yr | Yi] (MN-major, 128B
    //                  swizzle) into the (now free) input slot.  16-byte chunk index = j1.
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t re[8][4], im[8][4];
#pragma unroll
      for (int j2 = 0; j2 < 8; ++j2) {
        tmem_ld4(tD + 16 * j2 + 4 * c, re[j2]);
        tmem_ld4(tD + 16 * j2 + 8 + 4 * c, im[j2]);
      }
      tmem_ld_wait();
      reg_fence(re); reg_fence(im);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j1 = 4 * c + t;
        float vr[8], vi[8];
#pragma unroll
        for (int j2 = 0; j2 < 8; ++j2)
          cmul(__uint_as_float(re[j2][t]), __uint_as_float(im[j2][t]), twAr[j1], -twAi[j1], vr[j2], vi[j2]);
        const uint32_t off = uint32_t(lane) * 128u + (uint32_t(j1 ^ (lane & 7)) << 4);
        st_shared_v4(sX + off, pack_bf16x2(vr[0], vr[1]), pack_bf16x2(vr[2], vr[3]), pack_bf16x2(vr[4], vr[5]),
                     pack_bf16x2(vr[6], vr[7]));
        st_shared_v4(sX + kTileBytes + off, pack_bf16x2(vi[0], vi[1]), pack_bf16x2(vi[2], vi[3]),
                     pack_bf16x2(vi[4], vi[5]), pack_bf16x2(vi[6], vi[7]));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(bar_id, 128);
    // ---------------- stage 4: D4 = conj(F128) * Y   (lane = i, cols [0,64) -> seq b, [64,128) -> seq b+1)
    if (leader) {
      tc_fence_after();
      for (int s = 0; s < 8; ++s)
        mma_ts(tD0, tC0 + 8 * s, make_sdesc(sX + s * 2048, kTileBytes, 1024, 2), ID_N128_MN, s > 0);
      // D[:,0:64] += (-S) * Yi ;  D[:,64:128] += S * Yr       (conj F = C + iS)
      for (int s = 0; s < 8; ++s)
        mma_ts(tD0, tS0 + 8 * s, make_sdesc(sX + kTileBytes + s * 2048, kTileBytes, 1024, 2), ID_N64_MN_NEG, 1);
      for (int s = 0; s < 8; ++s)
        mma_ts(tD0 + 64, tS0 + 8 * s, make_sdesc(sX + s * 2048, kTileBytes, 1024, 2), ID_N64_MN, 1);
      mma_commit(bar_mma);
    }
    wait_mma();
    dump(first);

    // ---------------- pass 6: fp32 -> bf16, rows i of the two output tiles (128B swizzle), TMA store
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t re[32], im[32];
      tmem_ld32(tD + 32 * c, re);
      tmem_ld32(tD + 64 + 32 * c, im);
      tmem_ld_wait();
      reg_fence(re); reg_fence(im);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int cc = 4 * c + t;
        const uint32_t off = uint32_t(lane) * 128u + (uint32_t(cc ^ (lane & 7)) << 4);
        st_shared_v4(sX + off, pack_bf16x2(__uint_as_float(re[8 * t + 0]), __uint_as_float(re[8 * t + 1])),
                     pack_bf16x2(__uint_as_float(re[8 * t + 2]), __uint_as_float(re[8 * t + 3])),
                     pack_bf16x2(__uint_as_float(re[8 * t + 4]), __uint_as_float(re[8 * t + 5])),
                     pack_bf16x2(__uint_as_float(re[8 * t + 6]), __uint_as_float(re[8 * t + 7])));
        st_shared_v4(sX + kTileBytes + off,
                     pack_bf16x2(__uint_as_float(im[8 * t + 0]), __uint_as_float(im[8 * t + 1])),
                     pack_bf16x2(__uint_as_float(im[8 * t + 2]), __uint_as_float(im[8 * t + 3])),
                     pack_bf16x2(__uint_as_float(im[8 * t + 4]), __uint_as_float(im[8 * t + 5])),
                     pack_bf16x2(__uint_as_float(im[8 * t + 6]), __uint_as_float(im[8 * t + 7])));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(bar_id, 128);
    if (leader) {
      const int pr = unit - h * p.pairs;
      tma_store_3d(&tm_y, sX, 0, 0, seq_index(unit, 0));
      if (2 * pr + 1 < p.B) tma_store_3d(&tm_y, sX + kTileBytes, 0, 0, seq_index(unit, 1));
      tma_store_commit();
    }
  }

  // ---------------------------------------------------------------- teardown
  if (leader) tma_store_wait_all0();
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

}  // namespace r128
}  // namespace bffc

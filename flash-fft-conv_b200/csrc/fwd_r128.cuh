// Fused forward FFT-convolution kernel, N = 128 x 64 (= 8192), bf16, sm_100a.   (v2)
//
// Path replaced (reference): monarch_conv_cuda_kernel<32,8,8192,...>
// (csrc/flashfftconv/monarch_cuda/kernels_bf16/monarch_cuda_32_16_16_kernel_bf16.h:15-801) and its
// launcher (monarch_cuda_interface_fwd_bf16.cu:656-760).  Same math, different machine mapping:
//
//  * two real sequences (b, b+1) of one channel h are packed as ONE complex sequence z = u_b + i u_{b+1};
//    conv(z, k) = conv(u_b,k) + i conv(u_{b+1},k) because k is real, so no Hermitian split is needed.
//  * N = 128 * 64, n = i*64 + j.  Stage 1 contracts i: the 128x128 DFT matrix (cos / sin planes) is the
//    tcgen05 A operand and stays resident in TMEM for the whole kernel; the TMA-loaded (128 x 64) input tile
//    is the MN-major B operand.  D1[k1, j] lands in TMEM with lane = k1.
//  * stage 2 is "row local": lane k1 owns the 64-point transform over j.  The owning threads re-pack the
//    accumulator in TMEM as the next A operand (tcgen05.ld -> twiddle W_N^{k1 j} in fp32 registers ->
//    bf16x2 -> tcgen05.st) and one radix-64 complex MMA (4 real K=64 chains against the DFT-64 cos / sin
//    tiles in shared memory) produces Z[k1, k2], frequency k = k1 + 128*k2.
//  * pass 3 multiplies by k_f (pre-permuted to this order, "engine order"), stage 3 is the inverse radix-64.
//  * stage 4 contracts k1 (the lane index), so pass 5 writes its input to shared memory as the MN-major B
//    operand (same swizzled layout TMA produces); the output accumulator has lane = i; pass 6 converts to
//    bf16 and the tile is stored with TMA.
//  * a CTA runs two independent pipelines of 256 threads (two warpgroups each: the warpgroups split the
//    columns of every pass).  While one pipeline waits on its MMAs the other runs its CUDA-core pass.
//    No intermediate ever touches HBM.
#pragma once
#include "ptx.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace bffc {

struct FwdParams {
  const uint32_t* kf;        // [rows][16][128][4] bf16x2 words (kr0,kr1)(ki0,ki1)(kr2,kr3)(ki2,ki3), engine order, /N
  const __nv_bfloat16* dftC; // [128][128] cos(2*pi*m*k/128)
  const __nv_bfloat16* dftS; // [128][128] sin(2*pi*m*k/128)
  const uint8_t* gtiles;     // DFT-64 tiles Gr, Gi, -Gi, Gr: each 64 rows x 128 B, 128B-swizzled image
  float kf_scale;            // fp16 only: k_f is stored unscaled (1/N would underflow fp16) and scaled here in fp32
  float tw_scale;            // folded into the twiddle table (fp16: 1/sqrt(128) keeps every stage near the input level)
  const uint32_t* pregate;   // optional (B,H,L) bf16, or null
  const uint32_t* postgate;
  const uint32_t* postgate2; // optional second output gate: y2 = postgate2 * conv(...)  (gated backward: du and dpregate
  uint32_t* y2;              //   come from ONE pass, reference kernels_bf16/monarch_cuda_32_16_16_bwd_kernel_bf16.h:836-870)
  int B, H, L;               // batch, channels, sequence length
  int pairs;                 // ceil(B/2)
  int kmask;                 // bit s set: 16-row K step s of the input tile can be non-zero (the rest is skipped)
  int nseg;                  // segments per tile (small sizes: 4096/N batch members share one 8192 slot), else 1
  int seg_bytes;             // bytes of one segment inside a tile = (128 / nseg) rows x 128 B
  int small_out;             // 1: store the full tiles to the fold scratch, row = 2*unit + which
  int units;                 // H * pairs
  uint32_t kf_conj_mask;     // 0x80008000: multiply by conj(k_f) (du path of the backward: correlation), else 0
  float* dbg;                // optional stage dump [stage][128][128]
  int dbg_stages;
  long long* trace;          // bring-up builds only (-DBFFC_BRINGUP): clock64 stamps of CTA 0, [pipe][warp 0|3][unit][16]
};

namespace r128 {

constexpr int kThreads = 512;
constexpr int kPipeThreads = 256;
constexpr int kTileBytes = 128 * 128;          // one (128 rows x 64 bf16) tile
constexpr int kSlotBytes = 2 * kTileBytes;     // re tile + im tile
constexpr int kGTileBytes = 64 * 128;          // one DFT-64 plane
constexpr int kSmemData = 4 * kSlotBytes;      // 2 pipelines x 2 slots
constexpr int kSmemG = 4 * kGTileBytes;        // Gr, Gi, -Gi, Gr  (pairs at LBO 8K / 16K)
constexpr int kSmemBars = 64;
constexpr int kSmemGate = 2 * kSlotBytes;      // gated only: one pregate slot per pipeline
constexpr int kSmemTotal = kSmemData + kSmemG + kSmemBars + 1024;  // + alignment slack
constexpr int kSmemTotalGated = kSmemTotal + kSmemGate + 1024;

// TMEM columns
constexpr uint32_t kColC = 0, kColS = 64;                 // DFT-128 cos / sin, bf16 K-major A operand
DEVINL constexpr uint32_t colD(int pipe) { return 128 + 192 * pipe; }        // 128 fp32 cols
DEVINL constexpr uint32_t colA(int pipe) { return 128 + 192 * pipe + 128; }  // 64 cols (bf16x2)

template <int kFmt> struct Idesc {
  static constexpr uint32_t N128_MN = make_idesc(kFmt, 128, true, false);
  static constexpr uint32_t N64_MN = make_idesc(kFmt, 64, true, false);
  static constexpr uint32_t N64_MN_NEG = make_idesc(kFmt, 64, true, true);
};

DEVINL void cmul(float ar, float ai, float br, float bi, float& cr, float& ci) {
  cr = ar * br - ai * bi;
  ci = ar * bi + ai * br;
}
DEVINL uint64_t tile_desc(uint32_t saddr) { return make_sdesc(saddr, kTileBytes, 1024, 2); }

// One (128 x 64) input tile = nseg segments; segment s is the zero-padded (TMA out-of-bounds fill) start of batch member
// b = (g*nseg + s)*2 + which of channel h.  nseg == 1 is the ordinary case b = 2g + which.  A member beyond the batch
// is fetched from sequence index B*H, which is out of bounds for the tensor map: an all-zero tile.
DEVINL void load_tile(uint32_t dst, const void* map, uint32_t bar, int B, int H, int h, int g, int which, int nseg,
                      int seg_bytes) {
  for (int s = 0; s < nseg; ++s) {
    const int b = (g * nseg + s) * 2 + which;
    tma_load_3d(dst + s * seg_bytes, map, bar, 0, 0, b < B ? b * H + h : B * H);
  }
}
// N=128 B operand made of two 64-column tiles `lbo` bytes apart
DEVINL uint64_t pair_desc(uint32_t saddr, uint32_t lbo) { return make_sdesc(saddr, lbo, 1024, 2); }

DEVINL uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// kGated: y = postgate * conv(u * pregate, k)  (reference: GatedFlashFFTConvFunc, conv.py:3239-3325).  The
// pregate tiles arrive by TMA next to the input tiles and are multiplied in shared memory (bf16 product, as
// the reference's __hmul2 on load, monarch_cuda_32_16_16_kernel_bf16.h:550-585); the postgate is read with
// coalesced 16-byte loads in the output pass and applied to the bf16-rounded result.
// kPlanes: complex rows mode for composite sizes (N = R x 8192, see outer_cuda.cuh / outer_r128.cuh): the unit is one
// complex length-8192 row whose real / imaginary parts live in two bf16 planes (tm_u = tm_y = real plane,
// tm_g = imaginary plane), k_f row = unit / pairs (p.H = number of k_f rows), result written back in place.
template <bool kDebug, bool kGated, bool kPlanes = false, int kFmt = 1>
__global__ void __launch_bounds__(kThreads, 1)
fwd_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_y,
           const __grid_constant__ CUtensorMap tm_g, const FwdParams p) {
  using NT = Num<kFmt>;
  constexpr uint32_t ID_N128_MN = Idesc<kFmt>::N128_MN, ID_N64_MN = Idesc<kFmt>::N64_MN, ID_N64_MN_NEG = Idesc<kFmt>::N64_MN_NEG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_g = sbase + kSmemData;            // Gr tile, Gi tile
  const uint32_t s_bars = s_g + kSmemG;
  const uint32_t s_gate0 = s_bars + kSmemBars + 960;   // (gated only) keeps 1024-byte alignment
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));  // generic pointer to aligned base

  const int tid = threadIdx.x;
  const int pipe = __shfl_sync(0xffffffffu, tid >> 8, 0);   // pipeline = pair of warpgroups; warp-uniform for the compiler
  const int half = (tid >> 7) & 1;     // which 32-column half of every pass this warpgroup handles
  const int lane = tid & 127;          // TMEM lane owned by this thread (= k1, later = i)
  const int warp_q = (tid >> 5) & 3;   // TMEM sub-partition of this warp
  const bool lead_warp = ((tid & 255) < 32);   // warp that issues this pipeline's TMA / MMA (one elected lane)

  const uint32_t bar_tma0 = s_bars + pipe * 24;       // two TMA barriers
  const uint32_t bar_mma = s_bars + pipe * 24 + 16;   // one MMA barrier
  const uint32_t s_tmemptr = s_bars + 48;

  // ---------------------------------------------------------------- prologue
  if (tid == 0) {
    tma_prefetch_desc(&tm_u);
    tma_prefetch_desc(&tm_y);
    if (kGated || kPlanes) tma_prefetch_desc(&tm_g);
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
  {  // DFT-64 tiles -> smem (generic proxy writes, later read by the MMA/async proxy)
    const uint4* src = reinterpret_cast<const uint4*>(p.gtiles);
    uint4* dst = reinterpret_cast<uint4*>(gen_base + kSmemData);
    for (int i = tid; i < kSmemG / 16; i += kThreads) dst[i] = src[i];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + kSmemData + kSmemG + 48);
  const uint32_t tlane = tmem_base + (uint32_t(warp_q * 32) << 16);  // this warp's lane window

  // DFT-128 matrices -> TMEM: pipeline 0 loads cos, pipeline 1 loads sin; row = lane (128 bf16 = 64 cols),
  // each warpgroup of the pipeline loads 32 of the 64 columns.
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

  // lane dependent twiddles W_N^{k1*j} = c + i s, j = 32*half + 2q + {0,1}: kept as half2 pairs over two
  // adjacent points (|x| <= 1, 2^-12 relative) so the fp32x2 complex multiply can use them directly.
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
  const uint32_t sG0 = s_g;   // tiles: +0 Gr, +8K Gi, +16K -Gi, +24K Gr
  const f32x2 kfs2 = pk2(p.kf_scale, p.kf_scale);

  auto seq_index = [&](int unit, int which) {   // row of the output (and, in planes mode, input) tensor maps
    const int h = unit / p.pairs, pr = unit - h * p.pairs;
    if (kPlanes) return pr * p.H + h;           // row of both planes
    if (p.small_out) return 2 * unit + which;   // fold scratch
    return (2 * pr + which) * p.H + h;          // (b*H + h); members beyond the batch are not stored
  };
  auto issue_load = [&](int unit, int slot) {
    const uint32_t bar = bar_tma0 + 8 * slot;
    const uint32_t dst = s_slot0 + slot * kSlotBytes;
    const bool has_pre = kGated && p.pregate != nullptr;
    mbar_expect_tx(bar, has_pre ? 2 * kSlotBytes : kSlotBytes);
    const int uh = unit / p.pairs, ug = unit - uh * p.pairs;
    if (kPlanes) {
      tma_load_3d(dst, &tm_u, bar, 0, 0, seq_index(unit, 0));
      tma_load_3d(dst + kTileBytes, &tm_g, bar, 0, 0, seq_index(unit, 1));
    } else {
      load_tile(dst, &tm_u, bar, p.B, p.H, uh, ug, 0, p.nseg, p.seg_bytes);
      load_tile(dst + kTileBytes, &tm_u, bar, p.B, p.H, uh, ug, 1, p.nseg, p.seg_bytes);
    }
    if (has_pre) {   // single pregate slot per pipeline: free again once pass 0 of the current unit is done
      const uint32_t gd = s_gate0 + pipe * kSlotBytes;
      load_tile(gd, &tm_g, bar, p.B, p.H, uh, ug, 0, p.nseg, p.seg_bytes);
      load_tile(gd + kTileBytes, &tm_g, bar, p.B, p.H, uh, ug, 1, p.nseg, p.seg_bytes);
    }
  };

  uint32_t mma_phase = 0;
  auto wait_mma = [&]() {
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  };
  // all 256 threads of the pipeline finished writing TMEM (st) / smem; hand over to the MMA issuer
  auto sync_pipe_tmem = [&]() {
    tmem_st_wait();
    tc_fence_before();
    named_bar_sync(bar_id, kPipeThreads);
  };
  int dbg_stage = 0;
  auto dump = [&](bool first) {
    if (kDebug) {
      if (first && p.dbg != nullptr && dbg_stage < p.dbg_stages) {
        float* o = p.dbg + (size_t(dbg_stage) * 128 + lane) * 128;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tD + 64 * c + 32 * half, v);
          tmem_ld_wait();
          reg_fence(v);
#pragma unroll
          for (int t = 0; t < 32; ++t) o[64 * c + 32 * half + t] = __uint_as_float(v[t]);
        }
      }
      ++dbg_stage;
    }
  };
  if (lead_warp && u_begin < u_end) {
    if (elect_one()) issue_load(u_begin, 0);
    __syncwarp();
  }


  for (int unit = u_begin, n = 0; unit < u_end; ++unit, ++n) {
    const int slot = n & 1;
    const uint32_t sX = s_slot0 + slot * kSlotBytes;   // re tile; im tile at +kTileBytes
    const bool first = kDebug && (unit == 0);
    const int h = unit / p.pairs;

    if (kGated && p.pregate != nullptr) {
      // ---------------- pass 0: X <- bf16(u * pregate), in place in shared memory (same swizzled positions)
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
    // ---------------- stage 1: D1 = F128 * X   (lane = k1, cols [0,64) re, [64,128) im)
    if (lead_warp) {
      if (!(kGated && p.pregate != nullptr)) mbar_wait(bar_tma0 + 8 * slot, (n >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
      // D[:,0:128]  = C * [Xr | Xi];  D[:,0:64]  += S * Xi ;  D[:,64:128] += (-S) * Xr        (F = C - iS)
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
      }
      __syncwarp();
    }
    // k_f for pass 3: 32 packed complex of this (lane, half), coalesced 16 B per thread per chunk
    uint4 kfv[8];
    {
      const uint4* kfp = reinterpret_cast<const uint4*>(p.kf) + (size_t(h) * 16 + 8 * half) * 128 + lane;
#pragma unroll
      for (int c = 0; c < 8; ++c) kfv[c] = __ldg(kfp + c * 128);
    }
    wait_mma();
    dump(first);

    // ---------------- pass 1: * W^{k1*j} -> A1  (re part cols [0,32): K index = j, im part cols [32,64))
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
    sync_pipe_tmem();
    // ---------------- stage 2: radix-64 over j.  G = Gr + i Gi = exp(-2 pi i j k2 / 64)
    //   D[:,0:128] = re * [Gr | Gi] + im * [-Gi | Gr]
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dG0 = pair_desc(sG0, 8192), dG1 = pair_desc(sG0 + 16384, 8192);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ts(tD0, tA0 + 8 * s, dG0 + 128 * s, ID_N128_MN, s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ts(tD0, tA0 + 32 + 8 * s, dG1 + 128 * s, ID_N128_MN, 1);
        mma_commit(bar_mma);
        // prefetch the next unit into the other slot.  Its last reader was the previous unit's TMA store, issued
        // a full stage-1 + pass-1 ago: waiting for it here (not before stage 1) keeps the issuing warp from
        // arriving late at the pass-1 barrier (15 % barrier stall in profiles/r1_v21).
        if (unit + 1 < u_end) {
          tma_store_wait_read0();
          issue_load(unit + 1, slot ^ 1);
        }
      }
      __syncwarp();
    }
    wait_mma();
    dump(first);

    // ---------------- pass 3: * k_f  (frequency k = k1 + 128*k2) -> A3 (same layout as A1)
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
        const uint4 kq = kfv[4 * sub + (q >> 1)];
        const uint32_t wr = (q & 1) ? kq.z : kq.x, wi = ((q & 1) ? kq.w : kq.y) ^ p.kf_conj_mask;
        f32x2 vr, vi;
        f32x2 kr2 = NT::unpack(wr), ki2 = NT::unpack(wi);
        if (kFmt == 0) { kr2 = mul2(kr2, kfs2); ki2 = mul2(ki2, kfs2); }
        cmul2(pk2u(re[2 * q], re[2 * q + 1]), pk2u(im[2 * q], im[2 * q + 1]), kr2, ki2, vr, vi);
        ore[q] = NT::pack_v(vr);
        oim[q] = NT::pack_v(vi);
      }
      tmem_st8(tA + 16 * half + 8 * sub, ore);
      tmem_st8(tA + 32 + 16 * half + 8 * sub, oim);
    }
    sync_pipe_tmem();
    // ---------------- stage 3: inverse radix-64 (conj G):  D[:,0:128] = re * [Gr | -Gi] + im * [Gi | Gr]
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dG0 = pair_desc(sG0, 16384), dG1 = pair_desc(sG0 + 8192, 16384);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ts(tD0, tA0 + 8 * s, dG0 + 128 * s, ID_N128_MN, s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ts(tD0, tA0 + 32 + 8 * s, dG1 + 128 * s, ID_N128_MN, 1);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    wait_mma();
    dump(first);

    // ---------------- pass 5: * conj W^{k1*j}; write row k1 of the B operand [Yr | Yi] (MN-major, 128B swizzle)
    //                  into the (now free) input slot.  16-byte chunk index = j / 8.
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
        cmul2_conj(pk2u(re[2 * q], re[2 * q + 1]), pk2u(im[2 * q], im[2 * q + 1]), pk2(c.x, c.y), pk2(sn.x, sn.y), vr, vi);
        ore[q] = NT::pack_v(vr);
        oim[q] = NT::pack_v(vi);
      }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int chunk = 4 * half + 2 * sub + cc;
        const uint32_t off = uint32_t(lane) * 128u + (uint32_t(chunk ^ (lane & 7)) << 4);
        st_shared_v4(sX + off, ore[4 * cc + 0], ore[4 * cc + 1], ore[4 * cc + 2], ore[4 * cc + 3]);
        st_shared_v4(sX + kTileBytes + off, oim[4 * cc + 0], oim[4 * cc + 1], oim[4 * cc + 2], oim[4 * cc + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(bar_id, kPipeThreads);
    // ---------------- stage 4: D4 = conj(F128) * Y   (lane = i, cols [0,64) -> seq b, [64,128) -> seq b+1)
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dYr = tile_desc(sX), dYi = tile_desc(sX + kTileBytes);
#pragma unroll
        for (int s = 0; s < 8; ++s) mma_ts(tD0, tC0 + 8 * s, dYr + 128 * s, ID_N128_MN, s > 0);
        // D[:,0:64] += (-S) * Yi ;  D[:,64:128] += S * Yr       (conj F = C + iS)
#pragma unroll
        for (int s = 0; s < 8; ++s) mma_ts(tD0, tS0 + 8 * s, dYi + 128 * s, ID_N64_MN_NEG, 1);
#pragma unroll
        for (int s = 0; s < 8; ++s) mma_ts(tD0 + 64, tS0 + 8 * s, dYr + 128 * s, ID_N64_MN, 1);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    uint4 pg[2][4];
    const bool has_post = kGated && p.postgate != nullptr;
    if (has_post) {
      const bool row_ok = lane * 64 < p.L;
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        const int pb = 2 * (unit - h * p.pairs) + part;     // batch member (gated kernels never run segmented outputs)
        const uint4* gp_ = reinterpret_cast<const uint4*>(p.postgate + (size_t(pb < p.B ? pb : p.B - 1) * p.H + h) * p.L / 2 + lane * 32) + 4 * half;
#pragma unroll
        for (int c = 0; c < 4; ++c) pg[part][c] = row_ok ? __ldg(gp_ + c) : make_uint4(0, 0, 0, 0);
      }
    }
    wait_mma();
    dump(first);

    // ---------------- pass 6: fp32 -> bf16, rows i of the two output tiles (128B swizzle), TMA store
#pragma unroll
    for (int part = 0; part < 2; ++part) {          // 0: re -> sequence b, 1: im -> sequence b+1
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        uint32_t v[16];
        tmem_ld16(tD + 64 * part + 32 * half + 16 * sub, v);
        tmem_ld_wait();
        reg_fence(v);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int chunk = 4 * half + 2 * sub + cc;
          const uint32_t off = uint32_t(lane) * 128u + (uint32_t(chunk ^ (lane & 7)) << 4);
          uint32_t o0 = NT::pack(__uint_as_float(v[8 * cc + 0]), __uint_as_float(v[8 * cc + 1]));
          uint32_t o1 = NT::pack(__uint_as_float(v[8 * cc + 2]), __uint_as_float(v[8 * cc + 3]));
          uint32_t o2 = NT::pack(__uint_as_float(v[8 * cc + 4]), __uint_as_float(v[8 * cc + 5]));
          uint32_t o3 = NT::pack(__uint_as_float(v[8 * cc + 6]), __uint_as_float(v[8 * cc + 7]));
          if (kGated && p.y2 != nullptr) {
            // second gated output straight from registers: this thread owns 64 contiguous bytes of row `lane`
            const int pb = 2 * (unit - h * p.pairs) + part;
            if (pb < p.B && lane * 64 < p.L) {
              const size_t w = (size_t(pb) * p.H + h) * p.L / 2 + lane * 32 + 16 * half + 4 * (2 * sub + cc);
              const uint4 g2 = __ldg(reinterpret_cast<const uint4*>(p.postgate2 + w));
              *reinterpret_cast<uint4*>(p.y2 + w) =
                  make_uint4(NT::hmul2(o0, g2.x), NT::hmul2(o1, g2.y), NT::hmul2(o2, g2.z), NT::hmul2(o3, g2.w));
            }
          }
          if (has_post) {
            const uint4 g = pg[part][2 * sub + cc];
            o0 = NT::hmul2(o0, g.x); o1 = NT::hmul2(o1, g.y); o2 = NT::hmul2(o2, g.z); o3 = NT::hmul2(o3, g.w);
          }
          st_shared_v4(sX + part * kTileBytes + off, o0, o1, o2, o3);
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(bar_id, kPipeThreads);
    if (lead_warp) {
      if (elect_one()) {
        const int pr = unit - h * p.pairs;
        tma_store_3d(&tm_y, sX, 0, 0, seq_index(unit, 0));
        if (kPlanes) tma_store_3d(&tm_g, sX + kTileBytes, 0, 0, seq_index(unit, 1));
        else if (p.small_out || 2 * pr + 1 < p.B) tma_store_3d(&tm_y, sX + kTileBytes, 0, 0, seq_index(unit, 1));
        tma_store_commit();
      }
      __syncwarp();
    }
  }

  // ---------------------------------------------------------------- teardown
  if (lead_warp) tma_store_wait_all0();   // bulk groups are per thread: harmless on lanes that issued none
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

}  // namespace r128
}  // namespace bffc

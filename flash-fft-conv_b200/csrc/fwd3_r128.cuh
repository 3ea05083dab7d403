// Fused forward FFT-convolution kernel, N = 8192 — three pipelines of one warpgroup each; ungated, gated, and the
// complex-rows mode of the composite sizes.
//
// Same algorithm and stage list as r128_common.cuh.  The v2 kernel is latency-bound with only two sequence pairs in
// flight per SM (TMEM: 128 columns DFT-128 + 2 x (128 accumulator + 64 operand)); see profiles/r1_v21_summary.md.
// Here the A operand of the two radix-64 stages is staged in SHARED memory (the unit's own tile slot, K-major,
// 128B swizzle — byte-for-byte the layout pass 5 already writes for stage 4) instead of TMEM, which frees the
// operand columns: TMEM = 128 (DFT-128) + 3 x 128 accumulators, i.e. three pipelines of one warpgroup each.
// Stage 2 / 3 become SS-mode MMAs (A and B descriptors), stage 1 / 4 stay TS-mode (DFT-128 in TMEM).
//
// kGated: y = postgate * conv(u * pregate, k) (reference: GatedFlashFFTConvFunc, conv.py:3239-3325; __hmul2 on load /
// store, kernels_bf16/monarch_cuda_32_16_16_kernel_bf16.h:550-585).  Shared memory is full (3 x 2 slots + DFT-64 tiles),
// so a gated pipeline gives up the prefetch slot: slot 0 is the work area, slot 1 receives the gate tiles by TMA —
// the pregate next to the input (pass 0 multiplies in place, 16-bit product), then the postgate while the stages run
// (pass 6 multiplies the rounded result before the store), then optionally a second output gate (y2 = postgate2 * conv:
// du and dpregate of the gated backward from ONE pass, ...bwd_kernel_bf16.h:836-870; the accumulator is still in TMEM).
// The next unit's loads are issued when the last store has read slot 0; the other two pipelines cover that latency.
// p.xg_out: the gated input u * pregate is also stored (TMA, from slot 0 right after pass 0) — the backward's dk_f kernel
// consumes exactly these products, so the two gated passes of the backward hand them over instead of a separate
// elementwise pre-pass re-reading all four tensors.
#pragma once
#include "r128_common.cuh"

namespace bffc {
namespace r128 {

constexpr int kThreads3 = 384;
constexpr int kPipes3 = 3;
constexpr int kSmemData3 = kPipes3 * 2 * kSlotBytes;
constexpr int kSmemBars3 = 128;
constexpr int kSmemTotal3 = kSmemData3 + kSmemG + kSmemBars3 + 1024;

// K-major, 128B-swizzled A operand tile (128 rows x 64 bf16): 8-row groups 1024 B apart
DEVINL uint64_t atile_desc(uint32_t saddr) { return make_sdesc(saddr, 16, 1024, 2); }

struct GateMaps { CUtensorMap pre, post, post2, y2, xg; };

template <bool kPlanes, bool kGated, int kFmt>
__global__ void __launch_bounds__(kThreads3, 1)
fwd3_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_y,
            const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ GateMaps gm, const FwdParams p) {
  static_assert(!(kPlanes && kGated), "composite sizes apply their gates in the outer stages");
  using NT = Num<kFmt>;
  constexpr uint32_t ID_N128_MN = Idesc<kFmt>::N128_MN, ID_N64_MN = Idesc<kFmt>::N64_MN, ID_N64_MN_NEG = Idesc<kFmt>::N64_MN_NEG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_g = sbase + kSmemData3;
  const uint32_t s_bars = s_g + kSmemG;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));

  const int tid = threadIdx.x;
  // one warpgroup = one pipeline.  Taken through a shuffle so that the compiler knows it is warp-uniform: everything
  // the MMA issuer needs (slot addresses, TMEM columns, barriers, unit range) then lives in uniform registers.
  const int pipe = __shfl_sync(0xffffffffu, tid >> 7, 0);
  const int lane = tid & 127;          // TMEM lane (= k1, later = i)
  const int warp_q = (tid >> 5) & 3;
  const bool lead_warp = ((tid & 127) < 32);

  const uint32_t bar_tma0 = s_bars + pipe * 24;
  const uint32_t bar_mma = s_bars + pipe * 24 + 16;
  const uint32_t bar_g = s_bars + 80;
  const uint32_t s_tmemptr = s_bars + 96;

  if (tid == 0) {
    tma_prefetch_desc(&tm_u);
    tma_prefetch_desc(&tm_y);
    if (kPlanes) tma_prefetch_desc(&tm_g);
    if (kGated) {
      tma_prefetch_desc(&gm.pre); tma_prefetch_desc(&gm.post); tma_prefetch_desc(&gm.post2); tma_prefetch_desc(&gm.y2);
      tma_prefetch_desc(&gm.xg);
    }
  }
  if ((tid & 127) == 0) {
    mbar_init(bar_tma0, 1);
    mbar_init(bar_tma0 + 8, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (tid == 0) { mbar_init(bar_g, 1); fence_barrier_init(); }
  if (tid < 32) {
    tmem_alloc(s_tmemptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + kSmemData3 + kSmemG + 96);
  const uint32_t tlane = tmem_base + (uint32_t(warp_q * 32) << 16);

  const bool has_pre = kGated && p.pregate != nullptr, has_post = kGated && p.postgate != nullptr;
  const bool has_post2 = kGated && p.y2 != nullptr;
  const bool emit_xg = has_pre && p.xg_out != nullptr;
  const uint32_t bar_gate = bar_tma0 + 8;          // gated: slot 1 = gate tiles, its barrier counts postgate arrivals
  uint32_t gate_phase = 0;

  const int gp = blockIdx.x * kPipes3 + pipe;
  const int GP = gridDim.x * kPipes3;
  const int u_begin = int((long long)p.units * gp / GP);
  const int u_end = int((long long)p.units * (gp + 1) / GP);
  const uint32_t s_slot0 = sbase + pipe * 2 * kSlotBytes;

  auto seq_index = [&](int unit, int which) {      // complex-rows mode: plane row of the unit
    const int h = unit / p.pairs, pr = unit - h * p.pairs;
    (void)which;
    return pr * p.H + h;
  };
  auto issue_load = [&](int unit, int slot) {
    const uint32_t bar = bar_tma0 + 8 * slot;
    const uint32_t dst = s_slot0 + slot * kSlotBytes;
    if (kGated) {                        // input tiles -> slot 0, pregate tiles -> slot 1, one barrier
      const int uh = unit / p.pairs, ug = unit - uh * p.pairs;
      mbar_expect_tx(bar_tma0, has_pre ? 2 * kSlotBytes : kSlotBytes);
      load_tile(s_slot0, &tm_u, bar_tma0, p.B, p.H, uh, ug, 0, p.nseg, p.seg_bytes);
      load_tile(s_slot0 + kTileBytes, &tm_u, bar_tma0, p.B, p.H, uh, ug, 1, p.nseg, p.seg_bytes);
      if (has_pre) {
        load_tile(s_slot0 + kSlotBytes, &gm.pre, bar_tma0, p.B, p.H, uh, ug, 0, p.nseg, p.seg_bytes);
        load_tile(s_slot0 + kSlotBytes + kTileBytes, &gm.pre, bar_tma0, p.B, p.H, uh, ug, 1, p.nseg, p.seg_bytes);
      }
      return;
    }
    mbar_expect_tx(bar, kSlotBytes);
    if (kPlanes) {
      tma_load_3d(dst, &tm_u, bar, 0, 0, seq_index(unit, 0));
      tma_load_3d(dst + kTileBytes, &tm_g, bar, 0, 0, seq_index(unit, 1));
    } else {
      const int uh = unit / p.pairs, ug = unit - uh * p.pairs;
      load_tile(dst, &tm_u, bar, p.B, p.H, uh, ug, 0, p.nseg, p.seg_bytes);
      load_tile(dst + kTileBytes, &tm_u, bar, p.B, p.H, uh, ug, 1, p.nseg, p.seg_bytes);
    }
  };
  // Everything the first stage needs from global memory is requested up front and lands while the tables below are
  // built: the first unit's tiles (TMA), the DFT-64 tiles (one bulk copy, needed before the first stage 2 only).
  if (lead_warp && u_begin < u_end) {
    if (elect_one()) issue_load(u_begin, 0);
    __syncwarp();
  }
  if (tid == 0) {
    mbar_expect_tx(bar_g, kSmemG);
    for (int c = 0; c < kSmemG; c += 8192) bulk_load(s_g + c, reinterpret_cast<const uint8_t*>(p.gtiles) + c, 8192, bar_g);
  }

  // DFT-128 -> TMEM: pipeline 0 loads cos, pipeline 1 sin (64 columns each)
  if (pipe < 2) {
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
  // twiddles W_N^{k1*j} (k1 = lane; small sizes: lane mod N/64, period N), factored as A[j >> 3] * B[j & 7]: B (8 columns, half2 pairs) is a table, A is
  // advanced block by block with the per-lane step W_N^{8*k1} (fp32 recurrence over 8 blocks).  10 registers instead of
  // 64: with the full shared-memory carve-out there is no L1 behind local memory, so spills cost an L2 round trip.
  f32x2 twBc[4], twBs[4];
  float stc, sts;
  const int kl = lane & p.tw_mask;          // frequency index inside the stage-1 block (small sizes: N/64-point blocks)
  const float tw_inv = 1.0f / float(p.tw_n);
  sincospif(-2.0f * float(kl * 8) * tw_inv, &sts, &stc);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float s0, c0, s1, c1;
    sincospif(-2.0f * float(kl * (2 * q)) * tw_inv, &s0, &c0);
    sincospif(-2.0f * float(kl * (2 * q + 1)) * tw_inv, &s1, &c1);
    twBc[q] = pk2(c0, c1);
    twBs[q] = pk2(s0, s1);
  }
  // twiddles of the four column pairs of one 8-column block whose block factor is (ac, as)
  auto block_tw = [&](float ac, float as, f32x2 (&tc)[4], f32x2 (&ts)[4]) {
    const f32x2 ac2 = pk2(ac, ac), as2 = pk2(as, as);
#pragma unroll
    for (int q = 0; q < 4; ++q) cmul2(ac2, as2, twBc[q], twBs[q], tc[q], ts[q]);
  };
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t tD = tlane + 128 + 128 * pipe;
  const uint32_t tD0 = tmem_base + 128 + 128 * pipe;
  const uint32_t tC0 = tmem_base + kColC;
  const uint32_t tS0 = tmem_base + kColS;
  const uint32_t bar_id = 1 + pipe;
  const uint32_t sG0 = s_g;
  const f32x2 kfs2 = pk2(p.kf_scale, p.kf_scale);

#ifdef BFFC_BRINGUP
  // bring-up timeline (tools/trace_fwd3.py): lane 0 of warp 0 and of warp 3 of every pipeline of CTA 0
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && (tid & 31) == 0 && (warp_q == 0 || warp_q == 3);
  long long* trace_base = p.trace + (size_t(pipe) * 2 + (warp_q == 3)) * 64 * 16;
  int trace_n = 0;
  auto stamp = [&](int ev) { if (tracing && trace_n < 64) trace_base[trace_n * 16 + ev] = clock64(); };
#else
  auto stamp = [](int) {};
#endif
  uint32_t mma_phase = 0;
  auto wait_mma = [&]() {
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  };
  // hand the freshly written smem operand tiles to the MMA issuer
  auto sync_pipe_smem = [&]() {
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(bar_id, 128);
  };
  // row `lane` of the (re, im) tile pair: 16-byte chunk `chunk` (8 columns)
  auto store_chunk = [&](uint32_t sX, int chunk, const uint32_t* re4, const uint32_t* im4) {
    const uint32_t off = uint32_t(lane) * 128u + (uint32_t(chunk ^ (lane & 7)) << 4);
    st_shared_v4(sX + off, re4[0], re4[1], re4[2], re4[3]);
    st_shared_v4(sX + kTileBytes + off, im4[0], im4[1], im4[2], im4[3]);
  };

  // mirror of load_tile: the (up to two) tiles of a unit go back segment by segment, existing batch members only; rows
  // beyond L/64 of a segment are outside the tensor map and dropped
  auto store_tiles = [&](const CUtensorMap* my, uint32_t sT, int unit) {
    const int uh = unit / p.pairs, ug = unit - uh * p.pairs;
    for (int w = 0; w < 2; ++w)
      for (int sg = 0; sg < p.nseg; ++sg) {
        const int b = (ug * p.nseg + sg) * 2 + w;
        if (b < p.B) tma_store_3d(my, sT + w * kTileBytes + sg * p.seg_bytes, 0, 0, b * p.H + uh);
      }
    tma_store_commit();
  };

  if (lead_warp) mbar_wait(bar_g, 0);     // DFT-64 tiles have landed (long ago: hidden behind the table set-up)

  for (int unit = u_begin, n = 0; unit < u_end; ++unit, ++n) {
    const int slot = kGated ? 0 : (n & 1);
    const uint32_t sX = s_slot0 + slot * kSlotBytes;
    const uint32_t sGate = s_slot0 + kSlotBytes;
    const int h = unit / p.pairs;
    stamp(0);

    if (kGated) {
      // ---------------- pass 0: u * pregate in place (same swizzled image on both sides: linear 16-byte chunks)
      if (has_pre) {
        mbar_wait(bar_tma0, n & 1);
#pragma unroll 4
        for (int i = 0; i < kSlotBytes / 16 / 128; ++i) {
          const uint32_t off = uint32_t(i * 128 + lane) * 16u;
          const uint4 a = ld_shared_v4(sX + off), g = ld_shared_v4(sGate + off);
          st_shared_v4(sX + off, NT::hmul2(a.x, g.x), NT::hmul2(a.y, g.y), NT::hmul2(a.z, g.z), NT::hmul2(a.w, g.w));
        }
        sync_pipe_smem();                 // products visible to the tensor core; slot 1 is free again
      }
      if (emit_xg && lead_warp) {
        if (elect_one()) store_tiles(&gm.xg, sX, unit);
        __syncwarp();
      }
      if (has_post && lead_warp) {
        if (elect_one()) {
          const int ug = unit - h * p.pairs;
          mbar_expect_tx(bar_gate, kSlotBytes);
          load_tile(sGate, &gm.post, bar_gate, p.B, p.H, h, ug, 0, p.nseg, p.seg_bytes);
          load_tile(sGate + kTileBytes, &gm.post, bar_gate, p.B, p.H, h, ug, 1, p.nseg, p.seg_bytes);
        }
        __syncwarp();
      }
    }

    // ---------------- stage 1 (TS): D1 = F128 * X
    if (lead_warp) {
      mbar_wait(bar_tma0 + 8 * slot, kGated ? (n & 1) : ((n >> 1) & 1));
      tc_fence_after();
      stamp(1);
      if (elect_one()) {
        // descriptors are built once per unit; a K step moves the start-address field by (2048 >> 4)
        const uint64_t dXr = tile_desc(sX), dXi = tile_desc(sX + kTileBytes);
        if (p.kmask == 0xff) {      // full tiles: straight-line issue (the tensor pipe needs an MMA every 32-64 cycles)
#pragma unroll
          for (int s = 0; s < 8; ++s) mma_ts(tD0, tC0 + 8 * s, dXr + 128 * s, ID_N128_MN, s > 0);
#pragma unroll
          for (int s = 0; s < 8; ++s) mma_ts(tD0, tS0 + 8 * s, dXi + 128 * s, ID_N64_MN, 1);
#pragma unroll
          for (int s = 0; s < 8; ++s) mma_ts(tD0 + 64, tS0 + 8 * s, dXr + 128 * s, ID_N64_MN_NEG, 1);
        } else {                    // zero K steps (implicit padding, segmented small sizes) are skipped
          uint32_t acc = 0;
          for (int s = 0; s < 8; ++s)
            if ((p.kmask >> s) & 1) { mma_ts(tD0, tC0 + 8 * s, dXr + 128 * s, ID_N128_MN, acc); acc = 1; }
          for (int s = 0; s < 8; ++s)
            if ((p.kmask >> s) & 1) mma_ts(tD0, tS0 + 8 * s, dXi + 128 * s, ID_N64_MN, 1);
          for (int s = 0; s < 8; ++s)
            if ((p.kmask >> s) & 1) mma_ts(tD0 + 64, tS0 + 8 * s, dXr + 128 * s, ID_N64_MN_NEG, 1);
        }
        if (emit_xg) tma_store_wait_read0();   // pass 1 overwrites slot 0: nobody passes bar_mma before the store has read it
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    stamp(2);
    wait_mma();
    stamp(3);

    // ---------------- pass 1: * W^{k1 j} -> A1 tiles in the slot (K-major: row = lane, column = j)
    {
      float ac = p.tw_scale, as = 0.f;     // block factor W_N^{8*k1*block}
#pragma unroll 1
    for (int sub = 0; sub < 4; ++sub) {
      uint32_t re[16], im[16];
      tmem_ld16(tD + 16 * sub, re);
      tmem_ld16(tD + 64 + 16 * sub, im);
      tmem_ld_wait();
      reg_fence(re); reg_fence(im);
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        f32x2 tc[4], ts[4];
        block_tw(ac, as, tc, ts);
        uint32_t ore[4], oim[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          f32x2 vr, vi;
          cmul2(pk2u(re[8 * blk + 2 * q], re[8 * blk + 2 * q + 1]), pk2u(im[8 * blk + 2 * q], im[8 * blk + 2 * q + 1]), tc[q], ts[q], vr, vi);
          ore[q] = NT::pack_v(vr);
          oim[q] = NT::pack_v(vi);
        }
        store_chunk(sX, 2 * sub + blk, ore, oim);
        { const float nc = ac * stc - as * sts; as = ac * sts + as * stc; ac = nc; }
      }
    }
    }
    stamp(4);
    sync_pipe_smem();
    stamp(5);
    // ---------------- stage 2 (SS): D[:,0:128] = re * [Gr | Gi] + im * [-Gi | Gr]
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dAr = atile_desc(sX), dAi = atile_desc(sX + kTileBytes);
        const uint64_t dG0 = pair_desc(sG0, 8192), dG1 = pair_desc(sG0 + 16384, 8192);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ss(tD0, dAr + 2 * s, dG0 + 128 * s, ID_N128_MN, s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ss(tD0, dAi + 2 * s, dG1 + 128 * s, ID_N128_MN, 1);
        mma_commit(bar_mma);
        if (!kGated && unit + 1 < u_end) {
          tma_store_wait_read0();
          issue_load(unit + 1, slot ^ 1);
        }
      }
      __syncwarp();
    }
    // k_f of this lane: 16 vectors of 4 complex.  No L1 (see above): every load is an L2 round trip (~10 % of all stall
    // samples when fetched inside pass 3, profiles/r1_v3).  The first half is fetched here, where nothing else is live
    // (the MMA wait hides it), the second half at the start of pass 3 (hidden by the first half's arithmetic).
    const uint4* kfp = reinterpret_cast<const uint4*>(p.kf) + size_t(h) * 16 * 128 + lane;
    uint4 kfa[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) kfa[c] = __ldg(kfp + c * 128);
    stamp(6);
    wait_mma();
    stamp(7);

    // ---------------- pass 3: * k_f -> A3 tiles.  Each first-half vector is replaced by its second-half counterpart
    // as soon as it has been used (4 steps = several hundred cycles before the second half needs it).
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const int sub = 4 * hf + s4;
        uint32_t re[8], im[8];
        tmem_ld8(tD + 8 * sub, re);
        tmem_ld8(tD + 64 + 8 * sub, im);
        tmem_ld_wait();
        reg_fence(re); reg_fence(im);
        uint32_t ore[4], oim[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 kq = kfa[2 * s4 + (q >> 1)];
          const uint32_t wr = (q & 1) ? kq.z : kq.x, wi = ((q & 1) ? kq.w : kq.y) ^ p.kf_conj_mask;
          f32x2 kr2 = NT::unpack(wr), ki2 = NT::unpack(wi);
          if (kFmt == 0) { kr2 = mul2(kr2, kfs2); ki2 = mul2(ki2, kfs2); }
          f32x2 vr, vi;
          cmul2(pk2u(re[2 * q], re[2 * q + 1]), pk2u(im[2 * q], im[2 * q + 1]), kr2, ki2, vr, vi);
          ore[q] = NT::pack_v(vr);
          oim[q] = NT::pack_v(vi);
        }
        store_chunk(sX, sub, ore, oim);
        if (hf == 0) {
          kfa[2 * s4] = __ldg(kfp + (8 + 2 * s4) * 128);
          kfa[2 * s4 + 1] = __ldg(kfp + (9 + 2 * s4) * 128);
        }
      }
    }
    stamp(8);
    sync_pipe_smem();
    // ---------------- stage 3 (SS): inverse radix-64
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dAr = atile_desc(sX), dAi = atile_desc(sX + kTileBytes);
        const uint64_t dG0 = pair_desc(sG0, 16384), dG1 = pair_desc(sG0 + 8192, 16384);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ss(tD0, dAr + 2 * s, dG0 + 128 * s, ID_N128_MN, s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ss(tD0, dAi + 2 * s, dG1 + 128 * s, ID_N128_MN, 1);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    stamp(9);
    wait_mma();
    stamp(10);

    // ---------------- pass 5: * conj W -> Y tiles (MN-major B operand of stage 4; same bytes as a K-major A tile)
    {
      float ac = p.tw_scale, as = 0.f;     // block factor W_N^{8*k1*block}
#pragma unroll 1
    for (int sub = 0; sub < 4; ++sub) {
      uint32_t re[16], im[16];
      tmem_ld16(tD + 16 * sub, re);
      tmem_ld16(tD + 64 + 16 * sub, im);
      tmem_ld_wait();
      reg_fence(re); reg_fence(im);
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        f32x2 tc[4], ts[4];
        block_tw(ac, as, tc, ts);
        uint32_t ore[4], oim[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          f32x2 vr, vi;
          cmul2_conj(pk2u(re[8 * blk + 2 * q], re[8 * blk + 2 * q + 1]), pk2u(im[8 * blk + 2 * q], im[8 * blk + 2 * q + 1]), tc[q], ts[q], vr, vi);
          ore[q] = NT::pack_v(vr);
          oim[q] = NT::pack_v(vi);
        }
        store_chunk(sX, 2 * sub + blk, ore, oim);
        { const float nc = ac * stc - as * sts; as = ac * sts + as * stc; ac = nc; }
      }
    }
    }
    stamp(11);
    sync_pipe_smem();
    // ---------------- stage 4 (TS): conj F128 * Y
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dYr = tile_desc(sX), dYi = tile_desc(sX + kTileBytes);
#pragma unroll
        for (int s = 0; s < 8; ++s) mma_ts(tD0, tC0 + 8 * s, dYr + 128 * s, ID_N128_MN, s > 0);
#pragma unroll
        for (int s = 0; s < 8; ++s) mma_ts(tD0, tS0 + 8 * s, dYi + 128 * s, ID_N64_MN_NEG, 1);
#pragma unroll
        for (int s = 0; s < 8; ++s) mma_ts(tD0 + 64, tS0 + 8 * s, dYr + 128 * s, ID_N64_MN, 1);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    stamp(12);
    wait_mma();
    stamp(13);

    // ---------------- pass 6: fp32 -> 16 bit output tiles (x output gate), TMA store
    auto pass6 = [&](bool gate) {
#pragma unroll 1
      for (int sub = 0; sub < 4; ++sub) {
        uint32_t re[16], im[16];
        tmem_ld16(tD + 16 * sub, re);
        tmem_ld16(tD + 64 + 16 * sub, im);
        tmem_ld_wait();
        reg_fence(re); reg_fence(im);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          uint32_t ore[4], oim[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ore[q] = NT::pack(__uint_as_float(re[8 * blk + 2 * q]), __uint_as_float(re[8 * blk + 2 * q + 1]));
            oim[q] = NT::pack(__uint_as_float(im[8 * blk + 2 * q]), __uint_as_float(im[8 * blk + 2 * q + 1]));
          }
          if (kGated && gate) {
            const uint32_t off = uint32_t(lane) * 128u + (uint32_t((2 * sub + blk) ^ (lane & 7)) << 4);
            const uint4 g0 = ld_shared_v4(sGate + off), g1 = ld_shared_v4(sGate + kTileBytes + off);
            ore[0] = NT::hmul2(ore[0], g0.x); ore[1] = NT::hmul2(ore[1], g0.y); ore[2] = NT::hmul2(ore[2], g0.z); ore[3] = NT::hmul2(ore[3], g0.w);
            oim[0] = NT::hmul2(oim[0], g1.x); oim[1] = NT::hmul2(oim[1], g1.y); oim[2] = NT::hmul2(oim[2], g1.z); oim[3] = NT::hmul2(oim[3], g1.w);
          }
          store_chunk(sX, 2 * sub + blk, ore, oim);
        }
      }
    };
    auto store_out = [&](const CUtensorMap* my) {
      if (kPlanes) {
        tma_store_3d(my, sX, 0, 0, seq_index(unit, 0));
        tma_store_3d(&tm_g, sX + kTileBytes, 0, 0, seq_index(unit, 1));
        tma_store_commit();
      } else {
        store_tiles(my, sX, unit);
      }
    };
    if (has_post) { mbar_wait(bar_gate, gate_phase); gate_phase ^= 1; }
    pass6(has_post);
    stamp(14);
    sync_pipe_smem();
    if (lead_warp) {
      if (elect_one()) {
        store_out(&tm_y);
        if (has_post2) {                  // slot 1 has been read by every thread (barrier above): second gate -> slot 1
          const int ug = unit - h * p.pairs;
          mbar_expect_tx(bar_gate, kSlotBytes);
          load_tile(sGate, &gm.post2, bar_gate, p.B, p.H, h, ug, 0, p.nseg, p.seg_bytes);
          load_tile(sGate + kTileBytes, &gm.post2, bar_gate, p.B, p.H, h, ug, 1, p.nseg, p.seg_bytes);
          tma_store_wait_read0();         // the first output has left slot 0
        }
      }
      __syncwarp();
    }
    if (has_post2) {
      named_bar_sync(bar_id, 128);
      mbar_wait(bar_gate, gate_phase); gate_phase ^= 1;
      pass6(true);
      sync_pipe_smem();
      if (lead_warp) {
        if (elect_one()) store_out(&gm.y2);
        __syncwarp();
      }
    }
    if (kGated && lead_warp && unit + 1 < u_end) {
      if (elect_one()) {
        tma_store_wait_read0();
        issue_load(unit + 1, 0);
      }
      __syncwarp();
    }
    stamp(15);
#ifdef BFFC_BRINGUP
    ++trace_n;
#endif
  }

  if (lead_warp) tma_store_wait_all0();
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

}  // namespace r128
}  // namespace bffc

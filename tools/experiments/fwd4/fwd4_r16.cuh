// Fused forward FFT-convolution kernel, N = 8192 = 16 x 16 x 32 — the flop-lean three-radix variant.
//
// Path replaced (reference): monarch_conv_cuda_kernel<32,8,8192,...>, the same 32 x 16 x 16 class of factorisation
// (csrc/flashfftconv/monarch_cuda/kernels_bf16/monarch_cuda_32_16_16_kernel_bf16.h:590-763; tables conv.py:132-156),
// and, in "planes" mode, the complex inner Monarch convolution of the long sizes (monarch_cuda_*_complex_kernel_bf16.h).
//
// Why: the 128 x 64 kernel (fwd3_r128.cuh) issues 25.2 MFLOP of matmul per sequence pair — 3072 tensor-pipe cycles,
// 2x the reference's split — which both serialises its pipelines behind the tensor pipe and pins the board at its
// power cap.  Here every DFT stage takes the DATA as the tcgen05 A operand (M = 128 free positions) and a small DFT
// matrix as B (N = 2 x radix), so a stage costs 8 x radix cycles: 512 forward + 512 inverse = 8.4 MFLOP per pair, and
// no DFT matrix lives in TMEM.  What makes that possible without a transposing pass per stage:
//   * MN-major A operands.  The contracted index only has to be the ROW index of a shared-memory tile whose 128-byte
//     rows hold the free index.  A pass (thread = TMEM lane) then still writes 16-byte vectors of values it owns: the
//     digits it holds in its columns go to the fastest position of the next stage's M index, its lane digits to the
//     row / 16-byte-chunk position.  tests/kernel_model_r16.py states every layout and is checked against numpy.fft.
//   * The part of each inter-stage twiddle that depends on (output digit, column chunk) only is folded into the DFT
//     operand of the chunk; the rest is a 16-entry per-thread table rebuilt per pass by a short recurrence.
//   * The inverse chain ends with the top time digit in the columns, so its last two passes transpose: tcgen05.ld in
//     the 16x256b fragment layout + stmatrix.trans, which writes 8 consecutive lanes as one 16-byte row.
//   * Input and output tiles use the row order (n >> 6) & 7 major, n >> 9 minor (eight 16-row TMA boxes per plane), so
//     the first stage's K rows are adjacent and the transposing output pass is free of bank conflicts.
//
// Stage list per unit (one pair of sequences of one channel, or one complex row in planes mode):
//   S1 radix 16 over n[12:9]  ->  P1 * W_8192^{q1 m}          ->  S2 radix 16 over n''[8:5]  ->  P2 * W_512^{q2 c_lo}
//   S3 radix 32 over n[4:0]   ->  P3 * k_f (engine order v2)   ->  S3' inverse radix 32 (row local, K-major A)
//   P4 * conj W_8192^{c (q1 + 16 q2)}  ->  S2' inverse radix 16  ->  P5 (transposing) * conj W_256^{b q1}
//   S1' inverse radix 16      ->  P6 (transposing) 16-bit tiles ->  TMA store
// Machine mapping: one persistent CTA per SM, FOUR independent pipelines of one warpgroup (no DFT matrix lives in TMEM,
// so all 512 columns are accumulators: 4 x 128), and a ring of six 32 KB tile slots shared by them: the CTA's units are
// numbered j = 0, 1, ...; unit j runs on pipeline j % 4 in slot j % 6, so four slots are being worked on while two
// receive the next units.  A slot is handed from its previous user (unit j - 6, the partner pipeline (j + 2) % 4) to the
// pipeline that loads unit j through a "free" mbarrier, armed when the TMA store of unit j - 6 has read the slot.
#pragma once
#include "fwd3_r128.cuh"

namespace bffc {

struct Fwd4Params {
  const uint32_t* kf;        // engine order v2: [rows][16 vectors][128 lanes] x uint4 (re01, im01, re23, im23)
  const uint8_t* bmats;      // DFT operand images (kBmatBytes), see bffc.cu build_fwd4_tables()
  const float2* tw5;         // conj W_256^{b q1}, [q1][b]
  float kf_scale;            // fp16: scale applied with k_f in fp32 (k_f/N would underflow), else 1
  uint32_t kf_conj_mask;     // 0x80008000: multiply by conj(k_f)
  int B, H;
  int pairs;                 // ceil(B/2) (planes mode: complex rows per k_f row)
  int units;                 // H * pairs
  int planes;                // 1: unit = complex row `pr * H + h` of two planes
  int one_box;               // 1: the tensor maps deliver a whole plane (rows th-major) as ONE box; 0: eight boxes (one per th)
  float* dbg;                // bring-up: TMEM image after every stage [6][128][128] of unit 0
};

namespace r16 {

using namespace r128;       // tile geometry (kTileBytes, kSlotBytes), atile_desc

constexpr int kPipes4 = 4;
constexpr int kThreads4 = 128 * kPipes4;
constexpr int kSlots4 = 6;
constexpr int kSmemData4 = kSlots4 * kSlotBytes;
// DFT operand images, K-major without swizzle: [N rows][16 K] per K step, 8 x 16-byte core matrices
//   FWD16[chunk 0..3][plane]  4 x 2 x 1 KB   (W16^{k q} * W_64^{q chunk}: forward radix 16 + folded chunk twiddle)
//   INV16[plane]              2 x 1 KB
//   FWD32[plane][kstep]       2 x 2 x 2 KB
//   INV32[plane][kstep]       2 x 2 x 2 KB
constexpr int kOffF16 = 0, kOffI16 = 8192, kOffF32 = 10240, kOffI32 = 18432, kBmatBytes = 26624;
constexpr int kTw5Bytes = 16 * 16 * 8;
constexpr int kSmemBars4 = 192;
constexpr int kSmemTotal4 = kSmemData4 + kBmatBytes + kTw5Bytes + kSmemBars4 + 1024;

DEVINL constexpr uint32_t idesc_mn(int fmt, int n) {     // A MN-major, B K-major
  return (1u << 4) | (uint32_t(fmt) << 7) | (uint32_t(fmt) << 10) | (1u << 15) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24);
}
DEVINL constexpr uint32_t idesc_kk(int fmt, int n) {     // A K-major, B K-major
  return (1u << 4) | (uint32_t(fmt) << 7) | (uint32_t(fmt) << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24);
}
// MN-major 128B-swizzled A operand: 64-element atoms `lbo` bytes apart, 8-row K groups 1024 B apart
DEVINL uint64_t amn_desc(uint32_t saddr, uint32_t lbo) { return make_sdesc(saddr, lbo, 1024, 2); }
// K-major B operand without swizzle: the two 16-byte K halves 128 B apart, 8-row N groups 256 B apart
DEVINL uint64_t bk_desc(uint32_t saddr) { return make_sdesc(saddr, 128, 256, 0); }

DEVINL void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// four 8x8 16-bit matrices, transposed: register j of thread t holds M_j[2 (t % 4)][t / 4] (low half) and
// M_j[2 (t % 4) + 1][t / 4]; thread i supplies the address of the 16-byte row (i % 8) of matrix (i / 8)
DEVINL void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2),
               "r"(r3)
               : "memory");
}

// tm_*0 / tm_*1: 4-D maps [sequence][a = n >> 9][th = (n >> 6) & 7][64] of the two members of a pair (planes mode:
// real / imaginary plane); a box is the 16 rows (a) of one th.
template <bool kDebug, int kFmt>
__global__ void __launch_bounds__(kThreads4, 1)
fwd4_kernel(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1,
            const __grid_constant__ CUtensorMap tm_out0, const __grid_constant__ CUtensorMap tm_out1, const Fwd4Params p) {
  using NT = Num<kFmt>;
  constexpr uint32_t ID32 = idesc_mn(kFmt, 32), ID64 = idesc_mn(kFmt, 64), ID64K = idesc_kk(kFmt, 64);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_b = sbase + kSmemData4;
  const uint32_t s_tw5 = s_b + kBmatBytes;
  const uint32_t s_bars = s_tw5 + kTw5Bytes;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));

  const int tid = threadIdx.x;
  const int pipe = __shfl_sync(0xffffffffu, tid >> 7, 0);     // warp-uniform for the compiler (uniform-register MMA issue)
  const int lane = tid & 127;                                 // TMEM lane
  const int wq = (tid >> 5) & 3;                              // warp of the warpgroup = TMEM lane quarter
  const int wl = tid & 31;                                    // lane in warp
  const bool lead_warp = lane < 32;

  // barriers: full[slot] (TMA load landed), free[slot] (TMA store of the slot's previous unit has read it), mma[pipe]
  const uint32_t bar_full0 = s_bars, bar_free0 = s_bars + 48;
  const uint32_t bar_mma = s_bars + 96 + pipe * 8;
  const uint32_t bar_c = s_bars + 128;
  const uint32_t s_tmemptr = s_bars + 136;

  if (tid == 0) {
    tma_prefetch_desc(&tm_in0); tma_prefetch_desc(&tm_in1);
    tma_prefetch_desc(&tm_out0); tma_prefetch_desc(&tm_out1);
    mbar_init(bar_c, 1);
    for (int i = 0; i < kSlots4; ++i) { mbar_init(bar_full0 + 8 * i, 1); mbar_init(bar_free0 + 8 * i, 1); }
    for (int i = 0; i < kPipes4; ++i) mbar_init(s_bars + 96 + 8 * i, 1);
    fence_barrier_init();
  }
  if (tid < 32) {
    tmem_alloc(s_tmemptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + kSmemData4 + kBmatBytes + kTw5Bytes + 136);
  const uint32_t tlane = tmem_base + (uint32_t(wq * 32) << 16);

  // units of this CTA: [c_begin, c_begin + NU); local unit j -> pipeline j % kPipes4, slot j % kSlots4
  const int c_begin = int((long long)p.units * blockIdx.x / gridDim.x);
  const int NU = int((long long)p.units * (blockIdx.x + 1) / gridDim.x) - c_begin;

  // sequence (row) index of member `which` of a unit; beyond the batch: row B*H is out of bounds -> zero fill / dropped
  auto seq_index = [&](int unit, int which) {
    const int h = unit / p.pairs, pr = unit - h * p.pairs;
    if (p.planes) return pr * p.H + h;
    const int b = 2 * pr + which;
    return b < p.B ? b * p.H + h : p.B * p.H;
  };
  // TMA traffic of a pipeline is issued by one elected lane of its SECOND warp, so that the MMA-issuing lead warp never
  // sits behind it (one box per plane; eight when the driver rejects the th-major map, see make_map_r16)
  const bool tma_warp = (lane >> 5) == 1;
  auto issue_load = [&](int j) {            // local unit j into its slot
    const int unit = c_begin + j, slot = j % kSlots4;
    const uint32_t bar = bar_full0 + 8 * slot;
    const uint32_t dst = sbase + slot * kSlotBytes;
    mbar_expect_tx(bar, kSlotBytes);
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const CUtensorMap* tm = which ? &tm_in1 : &tm_in0;
      const int sq = seq_index(unit, which);
      if (p.one_box) tma_load_4d(dst + which * kTileBytes, tm, bar, 0, 0, 0, sq);
      else
        for (int th = 0; th < 8; ++th) tma_load_4d(dst + which * kTileBytes + th * 2048, tm, bar, 0, th, 0, sq);
    }
  };
  // all six slots start free: the first unit of this pipeline and, for the pipelines that own them, units kPipes4..5
  if (tma_warp) {
    if (elect_one()) {
      if (pipe < NU) issue_load(pipe);
      if (pipe + kPipes4 < kSlots4 && pipe + kPipes4 < NU) issue_load(pipe + kPipes4);
    }
    __syncwarp();
  }
  if (tid == 0) {
    mbar_expect_tx(bar_c, kBmatBytes + kTw5Bytes);
    for (int c = 0; c < kBmatBytes; c += 2048) bulk_load(s_b + c, p.bmats + c, 2048, bar_c);
    bulk_load(s_tw5, p.tw5, kTw5Bytes, bar_c);
  }

  // ---------------- per-thread twiddle seeds
  // P1: W_8192^{q1 m}, m = lane: pair (q1, q1 + 1) = pair(q1 - 2) * W^{2 m};  seeds W^m and W^{2m}
  float p1c, p1s, p1c2, p1s2;
  sincospif(-2.0f * float(lane) / 8192.0f, &p1s, &p1c);
  sincospif(-4.0f * float(lane) / 8192.0f, &p1s2, &p1c2);
  // P2: W_512^{q2 c_lo}, c_lo = (lane >> 3) & 7
  float p2c, p2s, p2c2, p2s2;
  {
    const int c_lo = (lane >> 3) & 7;
    sincospif(-2.0f * float(c_lo) / 512.0f, &p2s, &p2c);
    sincospif(-4.0f * float(c_lo) / 512.0f, &p2s2, &p2c2);
  }
  // P4: conj W_8192^{c s'}, s' = q1 + 16 q2 with q2 = 8 (lane >> 6) + (lane & 7), q1 = 8 q1_hi + ((lane >> 3) & 7)
  float p4c[2], p4s[2], p4c2[2], p4s2[2];
#pragma unroll
  for (int qh = 0; qh < 2; ++qh) {
    const int sp = 8 * qh + ((lane >> 3) & 7) + 16 * (8 * (lane >> 6) + (lane & 7));
    sincospif(2.0f * float(sp) / 8192.0f, &p4s[qh], &p4c[qh]);
    sincospif(4.0f * float(sp) / 8192.0f, &p4s2[qh], &p4c2[qh]);
  }

  const uint32_t tD = tlane + 128 * pipe;          // this warp's lane window of the pipeline's accumulator
  const uint32_t tD0 = tmem_base + 128 * pipe;
  const uint32_t bar_id = 1 + pipe;
  const f32x2 kfs2 = pk2(p.kf_scale, p.kf_scale);

  uint32_t mma_phase = 0;
  auto wait_mma = [&]() {
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  };
  auto sync_pipe_smem = [&]() {
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(bar_id, 128);
  };
  int dbg_stage = 0;
  auto dump = [&](bool first) {
    if (kDebug) {
      if (first && p.dbg != nullptr) {
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
  // radix-16 stage on MN-major tiles: 4 chunks x (re, im) MMAs of N = 32; B = forward (per chunk) or inverse operand
  auto issue_r16 = [&](uint32_t sX, bool inverse) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t bo = inverse ? s_b + kOffI16 : s_b + kOffF16 + c * 2048;
      mma_ss(tD0 + 32 * c, amn_desc(sX + 4096 * c, 2048), bk_desc(bo), ID32, 0);
      mma_ss(tD0 + 32 * c, amn_desc(sX + kTileBytes + 4096 * c, 2048), bk_desc(bo + 1024), ID32, 1);
    }
  };
  // 16-byte store of 8 consecutive M positions into 128-byte row `row` of both planes
  auto store_row_chunk = [&](uint32_t sX, int row, int chunk, const uint32_t* re4, const uint32_t* im4) {
    const uint32_t off = uint32_t(row) * 128u + (uint32_t(chunk ^ (row & 7)) << 4);
    st_shared_v4(sX + off, re4[0], re4[1], re4[2], re4[3]);
    st_shared_v4(sX + kTileBytes + off, im4[0], im4[1], im4[2], im4[3]);
  };

  if (lead_warp) mbar_wait(bar_c, 0);      // DFT operands have landed (hidden behind the seed set-up)

  for (int j = pipe; j < NU; j += kPipes4) {
    const int unit = c_begin + j, slot = j % kSlots4;
    const uint32_t sX = sbase + slot * kSlotBytes;
    const int h = unit / p.pairs;
    const bool first = kDebug && unit == 0;

    // ---------------- S1: radix 16 over the top time digit (raw tiles, rows (t, h, a))
    if (lead_warp) {
      mbar_wait(bar_full0 + 8 * slot, (j / kSlots4) & 1);
      tc_fence_after();
      if (elect_one()) {
        issue_r16(sX, false);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    wait_mma();
    dump(first);

    // ---------------- P1: * W_8192^{q1 m};  A2 row (2 c_hi + q1_hi) * 16 + b, chunk c_lo, b = 4 t + (m >> 5)
    {
      const int c_hi = (lane >> 3) & 3, c_lo = lane & 7, m_hi = lane >> 5;
      // table pairs (q1, q1 + 1), q1 = 0, 2, ..., 14
      f32x2 tc[8], ts[8];
      tc[0] = pk2(1.0f, p1c); ts[0] = pk2(0.0f, p1s);
      const f32x2 sc = pk2(p1c2, p1c2), ss = pk2(p1s2, p1s2);
#pragma unroll
      for (int j = 1; j < 8; ++j) cmul2(tc[j - 1], ts[j - 1], sc, ss, tc[j], ts[j]);
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        uint32_t re[16], im[16];
        tmem_ld16(tD + 32 * t, re);
        tmem_ld16(tD + 32 * t + 16, im);
        tmem_ld_wait();
        reg_fence(re); reg_fence(im);
        const int b = 4 * t + m_hi;
#pragma unroll
        for (int qh = 0; qh < 2; ++qh) {
          uint32_t ore[4], oim[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            f32x2 vr, vi;
            cmul2(pk2u(re[8 * qh + 2 * j], re[8 * qh + 2 * j + 1]), pk2u(im[8 * qh + 2 * j], im[8 * qh + 2 * j + 1]),
                  tc[4 * qh + j], ts[4 * qh + j], vr, vi);
            ore[j] = NT::pack_v(vr);
            oim[j] = NT::pack_v(vi);
          }
          store_row_chunk(sX, (2 * c_hi + qh) * 16 + b, c_lo, ore, oim);
        }
      }
    }
    sync_pipe_smem();
    // ---------------- S2: radix 16 over b
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        issue_r16(sX, false);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    wait_mma();
    dump(first);

    // ---------------- P2: * W_512^{q2 c_lo};  lane = 64 q1_hi + 8 c_lo + q1_lo;  A3 row (2 q1_hi + q2_hi) * 32 + c, chunk q1_lo
    {
      const int q1_hi = lane >> 6, c_lo = (lane >> 3) & 7, q1_lo = lane & 7;
      f32x2 tc[8], ts[8];
      tc[0] = pk2(1.0f, p2c); ts[0] = pk2(0.0f, p2s);
      const f32x2 sc = pk2(p2c2, p2c2), ss = pk2(p2s2, p2s2);
#pragma unroll
      for (int j = 1; j < 8; ++j) cmul2(tc[j - 1], ts[j - 1], sc, ss, tc[j], ts[j]);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t re[16], im[16];
        tmem_ld16(tD + 32 * ch, re);
        tmem_ld16(tD + 32 * ch + 16, im);
        tmem_ld_wait();
        reg_fence(re); reg_fence(im);
        const int c = 8 * ch + c_lo;
#pragma unroll
        for (int qh = 0; qh < 2; ++qh) {
          uint32_t ore[4], oim[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            f32x2 vr, vi;
            cmul2(pk2u(re[8 * qh + 2 * j], re[8 * qh + 2 * j + 1]), pk2u(im[8 * qh + 2 * j], im[8 * qh + 2 * j + 1]),
                  tc[4 * qh + j], ts[4 * qh + j], vr, vi);
            ore[j] = NT::pack_v(vr);
            oim[j] = NT::pack_v(vi);
          }
          store_row_chunk(sX, (2 * q1_hi + qh) * 32 + c, q1_lo, ore, oim);
        }
      }
    }
    sync_pipe_smem();
    // ---------------- S3: radix 32 over c (two chunks x (re, im) x two K steps, N = 64)
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              mma_ss(tD0 + 64 * t, amn_desc(sX + pl * kTileBytes + 8192 * t + 2048 * ks, 4096),
                     bk_desc(s_b + kOffF32 + (2 * pl + ks) * 2048), ID64, pl | ks);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    // prefetch this pipeline's next unit (half a unit of lead time covers HBM latency many times over)
    if (tma_warp) {
      const int jn = j + kPipes4;          // this pipeline's next unit (units below kSlots4 were loaded at the start)
      if (jn < NU && jn >= kSlots4) {
        if (elect_one()) {
          mbar_wait(bar_free0 + 8 * (jn % kSlots4), (jn / kSlots4 - 1) & 1);     // slot released by unit jn - 6
          issue_load(jn);
        }
        __syncwarp();
      }
    }
    // k_f of this lane: 16 vectors of 4 complex; the first half (chunk q1_hi = 0) is requested before the MMA wait
    const uint4* kfp = reinterpret_cast<const uint4*>(p.kf) + size_t(h) * 16 * 128 + lane;
    uint4 kfa[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) kfa[c] = __ldg(kfp + c * 128);
    wait_mma();
    dump(first);

    // ---------------- P3: * k_f;  K-major tile: row = lane, element 32 q1_hi + q3
#pragma unroll
    for (int qh = 0; qh < 2; ++qh) {
#pragma unroll
      for (int g2 = 0; g2 < 4; ++g2) {           // 8 values of q3 per step = vectors 2 g2, 2 g2 + 1 of this chunk
        uint32_t re[8], im[8];
        tmem_ld8(tD + 64 * qh + 8 * g2, re);
        tmem_ld8(tD + 64 * qh + 32 + 8 * g2, im);
        tmem_ld_wait();
        reg_fence(re); reg_fence(im);
        uint32_t ore[4], oim[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 kq = kfa[2 * g2 + (j >> 1)];
          const uint32_t wr = (j & 1) ? kq.z : kq.x, wi = ((j & 1) ? kq.w : kq.y) ^ p.kf_conj_mask;
          f32x2 kr2 = NT::unpack(wr), ki2 = NT::unpack(wi);
          if (kFmt == 0) { kr2 = mul2(kr2, kfs2); ki2 = mul2(ki2, kfs2); }
          f32x2 vr, vi;
          cmul2(pk2u(re[2 * j], re[2 * j + 1]), pk2u(im[2 * j], im[2 * j + 1]), kr2, ki2, vr, vi);
          ore[j] = NT::pack_v(vr);
          oim[j] = NT::pack_v(vi);
        }
        store_row_chunk(sX, lane, 4 * qh + g2, ore, oim);
        if (qh == 0) {
          kfa[2 * g2] = __ldg(kfp + (8 + 2 * g2) * 128);
          kfa[2 * g2 + 1] = __ldg(kfp + (9 + 2 * g2) * 128);
        }
      }
    }
    sync_pipe_smem();
    // ---------------- S3': inverse radix 32 over q3 (row local: K-major A, K steps of 32 bytes)
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              mma_ss(tD0 + 64 * t, atile_desc(sX + pl * kTileBytes + 64 * t + 32 * ks),
                     bk_desc(s_b + kOffI32 + (2 * pl + ks) * 2048), ID64K, pl | ks);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    wait_mma();
    dump(first);

    // ---------------- P4: * conj W_8192^{c (q1 + 16 q2)};  lane = 64 q2_hi + 8 q1_lo + q2_lo;
    //                  A2' row (2 c_hi + q1_hi) * 16 + q2, chunk q1_lo
    {
      const int q2 = 8 * (lane >> 6) + (lane & 7), q1_lo = (lane >> 3) & 7;
#pragma unroll 1
      for (int qh = 0; qh < 2; ++qh) {
        const float c1 = qh ? p4c[1] : p4c[0], s1 = qh ? p4s[1] : p4s[0], c2 = qh ? p4c2[1] : p4c2[0], s2 = qh ? p4s2[1] : p4s2[0];
        f32x2 wc = pk2(1.0f, c1), ws = pk2(0.0f, s1);       // pair (c, c + 1), advanced by W^{2 s'}
        const f32x2 sc = pk2(c2, c2), ss = pk2(s2, s2);
#pragma unroll
        for (int half = 0; half < 2; ++half) {       // 16 values of c per step
          uint32_t re[16], im[16];
          tmem_ld16(tD + 64 * qh + 16 * half, re);
          tmem_ld16(tD + 64 * qh + 32 + 16 * half, im);
          tmem_ld_wait();
          reg_fence(re); reg_fence(im);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {           // c_hi = 2 half + cc
            uint32_t ore[4], oim[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              f32x2 vr, vi;
              cmul2(pk2u(re[8 * cc + 2 * j], re[8 * cc + 2 * j + 1]), pk2u(im[8 * cc + 2 * j], im[8 * cc + 2 * j + 1]), wc, ws, vr, vi);
              ore[j] = NT::pack_v(vr);
              oim[j] = NT::pack_v(vi);
              f32x2 nc, ns;
              cmul2(wc, ws, sc, ss, nc, ns);
              wc = nc; ws = ns;
            }
            store_row_chunk(sX, (2 * (2 * half + cc) + qh) * 16 + q2, q1_lo, ore, oim);
          }
        }
      }
    }
    sync_pipe_smem();
    // ---------------- S2': inverse radix 16 over q2
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        issue_r16(sX, true);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    wait_mma();
    dump(first);

    // ---------------- P5 (transposing): * conj W_256^{b q1};  lane = 64 q1_hi + 8 q1_lo + c_lo, cols 32 c_hi + 16 ri + b;
    //                  A1' row (2 c_hi + b_hi) * 16 + q1, 16-byte chunk b_lo, element c_lo
    {
      const int b0 = 2 * (wl & 3);
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        // twiddle pairs (b0, b0 + 1) and (b0 + 8, b0 + 9) of the two lanes (j = 0, 1) of this thread
        f32x2 tc[2][2], ts[2][2];
        int q1v[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          q1v[j] = 8 * (wq >> 1) + 4 * (wq & 1) + 2 * hf + j;
#pragma unroll
          for (int gb = 0; gb < 2; ++gb) {
            const uint4 w = ld_shared_v4(s_tw5 + uint32_t(q1v[j] * 16 + 8 * gb + b0) * 8u);     // (c0, s0, c1, s1)
            tc[j][gb] = pk2u(w.x, w.z);
            ts[j][gb] = pk2u(w.y, w.w);
          }
        }
        // row address of this thread for stmatrix: thread i -> matrix i / 8 = (j, gb) = ((i >> 3) & 1, i >> 4), row r = i % 8
        const int mj = (wl >> 3) & 1, mgb = wl >> 4, mr = wl & 7;
        const int q1m = 8 * (wq >> 1) + 4 * (wq & 1) + 2 * hf + mj;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[16];
          tmem_ld_16x256b_x4(tD + (uint32_t(16 * hf) << 16) + 32 * ch, v);
          tmem_ld_wait();
          reg_fence(v);
          // v[4 g + 2 j + {0,1}]: lane j, columns 8 g + b0 + {0,1}; g = 0,1: re (b group gb = g), g = 2,3: im
          uint32_t ore[2][2], oim[2][2];
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int gb = 0; gb < 2; ++gb) {
              f32x2 vr, vi;
              cmul2(pk2u(v[4 * gb + 2 * j], v[4 * gb + 2 * j + 1]), pk2u(v[4 * (gb + 2) + 2 * j], v[4 * (gb + 2) + 2 * j + 1]),
                    tc[j][gb], ts[j][gb], vr, vi);
              ore[j][gb] = NT::pack_v(vr);
              oim[j][gb] = NT::pack_v(vi);
            }
          const int row = (2 * ch + mgb) * 16 + q1m;
          const uint32_t addr = sX + uint32_t(row) * 128u + (uint32_t(mr ^ (row & 7)) << 4);
          stmatrix_x4_trans(addr, ore[0][0], ore[1][0], ore[0][1], ore[1][1]);
          stmatrix_x4_trans(addr + kTileBytes, oim[0][0], oim[1][0], oim[0][1], oim[1][1]);
        }
      }
    }
    sync_pipe_smem();
    // ---------------- S1': inverse radix 16 over q1
    if (lead_warp) {
      tc_fence_after();
      if (elect_one()) {
        issue_r16(sX, true);
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    wait_mma();
    dump(first);

    // ---------------- P6 (transposing): lane = 64 b_hi + 8 b_lo + c_lo, cols 32 c_hi + 16 ri + a -> output tiles,
    //                  row 16 th + a, th = (b >> 1) & 7, 16-byte chunk 4 (b & 1) + c_hi, element c_lo
    {
      const int mj = (wl >> 3) & 1, mga = wl >> 4, mr = wl & 7;
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        const int bm = 8 * (wq >> 1) + 4 * (wq & 1) + 2 * hf + mj;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[16];
          tmem_ld_16x256b_x4(tD + (uint32_t(16 * hf) << 16) + 32 * ch, v);
          tmem_ld_wait();
          reg_fence(v);
          uint32_t ore[2][2], oim[2][2];
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int ga = 0; ga < 2; ++ga) {
              ore[j][ga] = NT::pack(__uint_as_float(v[4 * ga + 2 * j]), __uint_as_float(v[4 * ga + 2 * j + 1]));
              oim[j][ga] = NT::pack(__uint_as_float(v[4 * (ga + 2) + 2 * j]), __uint_as_float(v[4 * (ga + 2) + 2 * j + 1]));
            }
          const int row = 16 * ((bm >> 1) & 7) + 8 * mga + mr;
          const uint32_t addr = sX + uint32_t(row) * 128u + (uint32_t((4 * (bm & 1) + ch) ^ (row & 7)) << 4);
          stmatrix_x4_trans(addr, ore[0][0], ore[1][0], ore[0][1], ore[1][1]);
          stmatrix_x4_trans(addr + kTileBytes, oim[0][0], oim[1][0], oim[0][1], oim[1][1]);
        }
      }
    }
    sync_pipe_smem();
    if (tma_warp) {
      if (elect_one()) {
        const int pr = unit - h * p.pairs;
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
          if (!(p.planes || 2 * pr + which < p.B)) break;
          const CUtensorMap* tm = which ? &tm_out1 : &tm_out0;
          const int sq = seq_index(unit, which);
          if (p.one_box) tma_store_4d(tm, sX + which * kTileBytes, 0, 0, 0, sq);
          else
            for (int th = 0; th < 8; ++th) tma_store_4d(tm, sX + which * kTileBytes + th * 2048, 0, th, 0, sq);
        }
        tma_store_commit();
        // release the slot as soon as the store has read it (this lane would otherwise idle in the next unit's stage 1)
        tma_store_wait_read0();
        mbar_arrive(bar_free0 + 8 * slot);
      }
      __syncwarp();
    }
  }

  if (tma_warp) tma_store_wait_all0();
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

}  // namespace r16
}  // namespace bffc

// r128_common.cuh — definitions shared by the tcgen05 kernels built on the 8192 = 128 x 64 split (fwd3_r128.cuh,
// dkf3_r128.cuh, outer_r128.cuh): kernel parameter block, tile / slot geometry, TMEM column map, instruction and
// shared-memory descriptors, the segmented TMA tile load.
//
// Path replaced (reference): monarch_conv_cuda_kernel<32,8,8192,...>
// (csrc/flashfftconv/monarch_cuda/kernels_bf16/monarch_cuda_32_16_16_kernel_bf16.h:15-801) and its launcher
// (monarch_cuda_interface_fwd_bf16.cu:656-760).  Same math, different machine mapping:
//  * two real sequences (b, b+1) of one channel h are packed as ONE complex sequence z = u_b + i u_{b+1};
//    conv(z, k) = conv(u_b,k) + i conv(u_{b+1},k) because k is real, so no Hermitian split is needed.
//  * N = 128 * 64, n = i*64 + j.  Stage 1 contracts i: the 128x128 DFT matrix (cos / sin planes) is the tcgen05 A
//    operand and stays resident in TMEM for the whole kernel; the TMA-loaded (128 x 64) input tile is the MN-major B
//    operand.  D1[k1, j] lands in TMEM with lane = k1.  Stages 2 / 3 are radix-64 transforms over j against DFT-64
//    tiles resident in shared memory, stage 4 contracts k1 again; the CUDA-core passes between the MMAs apply the
//    twiddles and k_f ("engine order", frequency k = k1 + 128*k2).  No intermediate touches HBM.
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
  void* xg_out;              // gated three-pipeline kernel: also store u * pregate here (B,H,L), or null
  int B, H, L;               // batch, channels, sequence length
  int pairs;                 // ceil(B/2)
  int kmask;                 // bit s set: 16-row K step s of the input tile can be non-zero (the rest is skipped)
  int nseg;                  // segments per tile (small sizes: 8192/N batch members share one 8192-point slot), else 1
  int seg_bytes;             // bytes of one segment inside a tile = (128 / nseg) rows x 128 B
  int tw_n, tw_mask;         // stage-1 twiddles W_{tw_n}^{(lane & tw_mask) j}: 8192 / 127, small sizes N / (N/64 - 1)
  int units;                 // H * pairs
  uint32_t kf_conj_mask;     // 0x80008000: multiply by conj(k_f) (du path of the backward: correlation), else 0
  long long* trace;          // bring-up builds only (-DBFFC_BRINGUP): clock64 stamps of CTA 0, [pipe][warp 0|3][unit][16]
};

namespace r128 {

constexpr int kThreads = 512;                  // outer radix-128 stage: two pipelines of
constexpr int kPipeThreads = 256;              //   two warpgroups each
constexpr int kTileBytes = 128 * 128;          // one (128 rows x 64 bf16) tile
constexpr int kSlotBytes = 2 * kTileBytes;     // re tile + im tile
constexpr int kGTileBytes = 64 * 128;          // one DFT-64 plane
constexpr int kSmemG = 4 * kGTileBytes;        // Gr, Gi, -Gi, Gr  (pairs at LBO 8K / 16K)
constexpr int kSmemBars = 64;

// TMEM columns
constexpr uint32_t kColC = 0, kColS = 64;                 // DFT-128 cos / sin, bf16 K-major A operand
DEVINL constexpr uint32_t colD(int pipe) { return 128 + 192 * pipe; }        // outer stage: 128 fp32 accumulator columns

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

// One (128 x 64) input tile = nseg segments of 128/nseg rows; segment s holds batch member b = (g*nseg + s)*2 + which
// of channel h (rows beyond L/64: TMA out-of-bounds zero fill = implicit padding).  nseg == 1 is the ordinary case
// b = 2g + which.  Small sizes N < 8192: nseg = 8192/N members, each an independent N-point circular convolution —
// stage 1 uses the block-diagonal matrix I_nseg (x) F_{N/64} instead of F_128, the rest of the kernel is unchanged.
// A member beyond the batch is fetched from sequence index B*H, out of bounds for the tensor map: an all-zero tile.
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

}  // namespace r128
}  // namespace bffc

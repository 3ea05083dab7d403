// bffc.cu — host side of the C ABI declared in include/bffc.h (plan tables, TMA descriptors, launches).
//
// Replaces, for the fused FFT-convolution path only:
//   FlashFFTConv.__init__ tables            (reference flashfftconv/conv.py:72-551)
//   k_f permutation + cast per call         (conv.py:640, :676, :1423-1424)
//   pybind entry + C++ dispatch + launcher  (csrc/flashfftconv/monarch.cpp:16-56,
//                                            monarch_cuda/monarch_fwd.h:296-376,
//                                            monarch_cuda_interface_fwd_bf16.cu:656-760)
//   the three-kernel orchestration of the long sizes (conv.py:1420-1524: butterfly -> complex Monarch -> ibutterfly)
// No torch types cross this boundary; there is no CPU fallback.
#include "bffc.h"
#include "r128_common.cuh"
#include "fwd3_r128.cuh"
#include "dkf3_r128.cuh"
#include "outer_cuda.cuh"
#include "outer_r128.cuh"
#include "filter_fft.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

thread_local char g_err[512] = "";
thread_local int g_launches = 0;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) return fail(BFFC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;
std::once_flag g_encode_once;

int get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres) ==
            cudaSuccess && qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  });
  return g_encode ? 0 : fail(BFFC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
}

int check_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(BFFC_ERR_NO_DEVICE, "no CUDA device available (bffc has no CPU fallback)");
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
    cudaGetLastError();
    return fail(BFFC_ERR_NO_DEVICE, "device %d is not sm_100 (compute capability major %d)", dev, major);
  }
  return 0;
}

// column-FFT kernels of the filter side are instantiated per outer size R = N / 8192
#define COLS_SWITCH(R_, ...)                                                                          \
  do {                                                                                               \
    switch (R_) {                                                                                    \
      case 2: { constexpr int RR = 2; __VA_ARGS__ } break;                                           \
      case 4: { constexpr int RR = 4; __VA_ARGS__ } break;                                           \
      case 8: { constexpr int RR = 8; __VA_ARGS__ } break;                                           \
      case 16: { constexpr int RR = 16; __VA_ARGS__ } break;                                         \
      case 32: { constexpr int RR = 32; __VA_ARGS__ } break;                                         \
      case 64: { constexpr int RR = 64; __VA_ARGS__ } break;                                         \
      case 128: { constexpr int RR = 128; __VA_ARGS__ } break;                                       \
      case 256: { constexpr int RR = 256; __VA_ARGS__ } break;                                       \
      default: { constexpr int RR = 512; __VA_ARGS__ } break;                                        \
    }                                                                                                \
  } while (0)
#define FMT_SWITCH(dtype_, ...)                         \
  do {                                                   \
    if ((dtype_) == BFFC_DTYPE_BF16) { constexpr int F = 1; __VA_ARGS__ } \
    else { constexpr int F = 0; __VA_ARGS__ }            \
  } while (0)

uint16_t f2bf(double x);
uint16_t f2h16(double x, int dtype) {   // table entry in the plan's element type
  if (dtype == BFFC_DTYPE_BF16) return f2bf(x);
  __half h = __float2half_rn(static_cast<float>(x));
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}

uint16_t f2bf(double x) {  // round-to-nearest-even float -> bf16 bits
  float f = static_cast<float>(x);
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t r = 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>((u + r) >> 16);
}

// k_f -> engine order.  One thread produces one 16-byte engine vector = 4 consecutive inner frequencies k2 = 4cc..4cc+3
// of one (row, k1): words (re k2, re k2+1) (im ..) (re k2+2, re k2+3) (im ..).  Reads are coalesced along k1
// (stride R complex numbers), writes are fully coalesced.
// kHalf: the source holds only frequencies 0..N/2 of a real filter (torch.fft.rfft); k > N/2 is conj(src[N-k]).
template <bool kHalf, int kFmt>
__global__ void kf_pack_kernel(const float2* __restrict__ kf_nat, uint4* __restrict__ kf_eng, int N, int R0, int R1,
                               float scale, int conj, int rblk) {
  const int h = blockIdx.y;
  const int R = R0 * R1;
  const float2* src = kf_nat + size_t(h) * (kHalf ? (N / 2 + 1) : N);
  const int nvec = N / 4;                              // engine vectors per channel
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
    const int row = v / 2048, rem = v % 2048;          // 2048 vectors per 8192-word row
    const int cc = rem >> 7, k1 = rem & 127;
    const int c0 = row / R1, c1 = row % R1;
    float2 e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // small sizes (rblk = seqlen/64 < 128): K_seqlen[f] = K_8192[f * 128/rblk] at f = (k1 mod rblk) + rblk k2
      int k = c0 + R0 * (c1 + R1 * (((k1 & (rblk - 1)) + rblk * (4 * cc + i)) * (128 / rblk)));
      float sg = conj ? -1.f : 1.f;
      if (kHalf && k > N / 2) { k = N - k; sg = -sg; }
      const float2 t = src[k];
      e[i] = make_float2(t.x * scale, t.y * scale * sg);
    }
    using NT = bffc::Num<kFmt>;
    kf_eng[size_t(h) * nvec + v] = make_uint4(NT::pack(e[0].x, e[1].x), NT::pack(e[0].y, e[1].y),
                                              NT::pack(e[2].x, e[3].x), NT::pack(e[2].y, e[3].y));
  }
  (void)R;
}

// Tiled variant for a large outermost radix R0 (tcgen05 outer stage, R0 = 128): consecutive c0 are adjacent in the
// natural order, consecutive words are adjacent in the engine rows, so a (32 c0) x (32 word pairs) tile goes
// through shared memory and both sides are accessed in 256-byte runs.
// grid: (8192/2/32 word-pair tiles, R0/32 * R1, H)
constexpr int kInnerWords = 8192;

template <bool kHalf, int kFmt>
__global__ void kf_pack_tiled_kernel(const float2* __restrict__ kf_nat, uint2* __restrict__ kf_eng, int N, int R0, int R1,
                                     float scale, int conj) {
  __shared__ uint2 tile[32][33];
  const int h = blockIdx.z;
  const int c0b = (blockIdx.y % (R0 / 32)) * 32, c1 = blockIdx.y / (R0 / 32);
  const int wp0 = blockIdx.x * 32;
  const int R = R0 * R1;
  const float2* src = kf_nat + size_t(h) * (kHalf ? (N / 2 + 1) : N);
  const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int wp = wp0 + j;                            // word pair index inside the row: (cc*128 + k1)*2 + pp
    const int pp = wp & 1, k1 = (wp >> 1) & 127, cc = wp >> 8;
    float2 v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      int k = (c0b + tx) + R0 * (c1 + R1 * (k1 + 128 * (4 * cc + 2 * pp + e)));
      float sg = conj ? -1.f : 1.f;
      if (kHalf && k > N / 2) { k = N - k; sg = -sg; }
      float2 t = src[k];
      v[e] = make_float2(t.x * scale, t.y * scale * sg);
    }
    tile[tx][j] = make_uint2(bffc::Num<kFmt>::pack(v[0].x, v[1].x), bffc::Num<kFmt>::pack(v[0].y, v[1].y));
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {                   // j = c0 offset, tx = word pair
    const size_t row = (size_t(h) * R0 + (c0b + j)) * R1 + c1;
    kf_eng[row * (kInnerWords / 2) + wp0 + tx] = tile[j][tx];
  }
  (void)R;
}

constexpr int kInner = 8192;   // the fused tcgen05 kernel's size

}  // namespace

struct bffc_level { int tc; int R; };   // tc = 1: tcgen05 radix-128 stage (outer_r128.cuh); 0: CUDA-core radix 2/4/8

struct bffc_plan {
  int NE;        // engine FFT size: N for N >= 8192; 8192 for the small sizes (256..4096): 8192/N batch members of a
                 // channel run as independent N-point circular convolutions inside one 8192-point unit (stage 1 =
                 // block-diagonal I (x) F_{N/64}), see r128_common.cuh
  int N;
  int R;         // N = R * 8192: product of the outer radices around the fused 8192-point kernel
  int nlev;      // number of outer levels (0, 1 or 2), outermost first
  bffc_level lev[2];
  int dtype;
  int device;
  __nv_bfloat16* dftC = nullptr;
  __nv_bfloat16* dftS = nullptr;
  uint8_t* gtiles = nullptr;
  float2* tw8192 = nullptr;   // e^{-2 pi i t / 8192}, t < 8192: twiddles of the fp32 filter-side FFTs (filter_fft.cuh)
  float2* tw512 = nullptr;    // composite sizes: e^{-2 pi i t / 512} (column FFTs), W_N^{j} j < 2048, W_N^{2048 i} i < N/2048
  float2* tw_lo = nullptr;
  float2* tw_hi = nullptr;
  int num_sms = 0;
  // bffc_fwd_host: copy-in / compute / copy-out streams and the per-slot events (created with the plan)
  cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t hev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

extern "C" {

int bffc_abi_version(void) { return BFFC_ABI_VERSION; }
const char* bffc_last_error(void) { return g_err; }
int bffc_last_launch_count(void) { return g_launches; }

static int levels_for(int N, bffc_level* lev) {
  switch (N) {
    case 256: case 512: case 1024: case 2048: case 4096: return 0;
    case 8192: return 0;
    case 16384: lev[0] = {0, 2}; return 1;
    case 32768: lev[0] = {0, 4}; return 1;
    case 65536: lev[0] = {0, 8}; return 1;
    case 131072: lev[0] = {0, 8}; lev[1] = {0, 2}; return 2;
    case 262144: lev[0] = {0, 8}; lev[1] = {0, 4}; return 2;
    case 524288: lev[0] = {0, 8}; lev[1] = {0, 8}; return 2;
    case 1048576: lev[0] = {1, 128}; return 1;
    case 2097152: lev[0] = {1, 128}; lev[1] = {0, 2}; return 2;
    case 4194304: lev[0] = {1, 128}; lev[1] = {0, 4}; return 2;
    default: return -1;
  }
}

int bffc_plan_destroy(bffc_plan* p);

int bffc_supported(int seqlen, int dtype) {
  if (dtype != BFFC_DTYPE_BF16 && dtype != BFFC_DTYPE_FP16) return 0;
  bffc_level lev[2];
  return levels_for(seqlen, lev) >= 0 ? 1 : 0;
}

int bffc_plan_create(bffc_plan** out, int seqlen, int dtype) {
  if (!out) return fail(BFFC_ERR_INVALID, "plan output pointer is null");
  *out = nullptr;
  if (dtype != BFFC_DTYPE_BF16 && dtype != BFFC_DTYPE_FP16) return fail(BFFC_ERR_INVALID, "unknown dtype %d", dtype);
  if (!bffc_supported(seqlen, dtype))
    return fail(BFFC_ERR_UNSUPPORTED, "seqlen %d / dtype %d not supported by this build", seqlen, dtype);
  if (int rc = check_device()) return rc;
  if (int rc = get_encode()) return rc;

  bffc_plan* p = new bffc_plan();
  // every failure below releases what was created so far (bffc_plan_destroy accepts a partially built plan)
#define PLAN_TRY(expr)                                                                               \
  do {                                                                                               \
    cudaError_t e_ = (expr);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      bffc_plan_destroy(p);                                                                          \
      return fail(BFFC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));                    \
    }                                                                                                \
  } while (0)
  p->N = seqlen;
  p->NE = seqlen < kInner ? kInner : seqlen;
  p->R = p->NE / kInner;
  p->nlev = levels_for(seqlen, p->lev);
  p->dtype = dtype;
  PLAN_TRY(cudaGetDevice(&p->device));
  PLAN_TRY(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));

  const double PI = 3.14159265358979323846;
  // stage-1 DFT of the fused kernel, cos / sin planes (symmetric, K-major rows): F_128, or for the small sizes the
  // block-diagonal I_{8192/N} (x) F_r, r = N/64 (one block per batch member sharing the unit)
  std::vector<uint16_t> c(128 * 128), s(128 * 128);
  const int rblk = seqlen < kInner ? seqlen / 64 : 128;
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 128; ++k) {
      const bool same = m / rblk == k / rblk;
      const double ang = 2.0 * PI * double(((m % rblk) * (k % rblk)) % rblk) / double(rblk);
      c[m * 128 + k] = f2h16(same ? cos(ang) : 0.0, dtype);
      s[m * 128 + k] = f2h16(same ? sin(ang) : 0.0, dtype);
    }
  PLAN_TRY(cudaMalloc(&p->dftC, c.size() * 2));
  PLAN_TRY(cudaMalloc(&p->dftS, s.size() * 2));
  PLAN_TRY(cudaMemcpy(p->dftC, c.data(), c.size() * 2, cudaMemcpyHostToDevice));
  PLAN_TRY(cudaMemcpy(p->dftS, s.data(), s.size() * 2, cudaMemcpyHostToDevice));

  // DFT-64 planes for the row-local stage: G = exp(-2 pi i k n / 64) = Gr + i Gi.  Stored as the MN-major
  // B operand image: row k (K index) = 64 bf16 = 128 B, 16-byte chunk c of row k at chunk position c ^ (k & 7)
  // (the 128B swizzle TMA / UMMA use).
  std::vector<uint8_t> gt(4 * bffc::r128::kGTileBytes, 0);   // tiles: Gr, Gi, -Gi, Gr
  for (int k = 0; k < 64; ++k)
    for (int n = 0; n < 64; ++n) {
      const double ang = -2.0 * PI * double((k * n) & 63) / 64.0;
      const uint16_t gr = f2h16(cos(ang), dtype), gi = f2h16(sin(ang), dtype), ngi = f2h16(-sin(ang), dtype);
      const size_t off = size_t(k) * 128 + (size_t((n >> 3) ^ (k & 7)) << 4) + size_t(n & 7) * 2;
      const size_t T = bffc::r128::kGTileBytes;
      memcpy(gt.data() + off, &gr, 2);
      memcpy(gt.data() + T + off, &gi, 2);
      memcpy(gt.data() + 2 * T + off, &ngi, 2);
      memcpy(gt.data() + 3 * T + off, &gr, 2);
    }
  PLAN_TRY(cudaMalloc(&p->gtiles, gt.size()));
  PLAN_TRY(cudaMemcpy(p->gtiles, gt.data(), gt.size(), cudaMemcpyHostToDevice));

  {
    auto table = [&](float2** dst, int n, double period) -> cudaError_t {
      std::vector<float2> tw(n);
      for (int t = 0; t < n; ++t) {
        const double a = -2.0 * M_PI * double(t) / period;
        tw[t] = make_float2(float(cos(a)), float(sin(a)));
      }
      cudaError_t e = cudaMalloc(dst, tw.size() * sizeof(float2));
      return e != cudaSuccess ? e : cudaMemcpy(*dst, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice);
    };
    PLAN_TRY(table(&p->tw8192, kInner, double(kInner)));
    if (p->NE > kInner) {
      PLAN_TRY(table(&p->tw512, 512, 512.0));
      PLAN_TRY(table(&p->tw_lo, 2048, double(p->NE)));
      PLAN_TRY(table(&p->tw_hi, p->NE / 2048, double(p->NE) / 2048.0));
    }
  }

  // engine order of k_f (32-bit words), per channel h: R rows of 8192 words, row = c0*R1 + c1 (outer digits); inside a
  // row  w = (cc*128 + k1)*4 + 2*pp + part  holds the 16-bit pair (part ? imag : real) of k_f at inner frequencies
  // k'' = k1 + 128*k2 with k2 = 4cc + 2pp and k2 + 1; natural frequency k = c0 + R0*(c1 + R1*k'').  See kf_pack_kernel.

  using namespace bffc::r128;
  FMT_SWITCH(dtype,
    PLAN_TRY(cudaFuncSetAttribute(fwd3_kernel<false, false, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal3));
    PLAN_TRY(cudaFuncSetAttribute(fwd3_kernel<false, true, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal3));
    PLAN_TRY(cudaFuncSetAttribute(fwd3_kernel<true, false, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal3));
    PLAN_TRY(cudaFuncSetAttribute(dkf3_kernel<true, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotalDkf3));
    PLAN_TRY(cudaFuncSetAttribute(dkf3_kernel<false, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotalDkf3));
    PLAN_TRY(cudaFuncSetAttribute(outer_tc_kernel<false, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOuterGated));
    PLAN_TRY(cudaFuncSetAttribute(outer_tc_kernel<true, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOuter));
    PLAN_TRY(cudaFuncSetAttribute(bffc::ffft::kf_from_filter_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, bffc::ffft::kSmemBytes));
  );
  PLAN_TRY(cudaFuncSetAttribute(bffc::ffft::dk_from_dkf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bffc::ffft::kSmemBytes));
  if (p->NE > kInner) {
    using namespace bffc::ffft;
    FMT_SWITCH(dtype, PLAN_TRY(cudaFuncSetAttribute(filter_rows_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)););
    PLAN_TRY(cudaFuncSetAttribute(dk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    COLS_SWITCH(p->R, PLAN_TRY(cudaFuncSetAttribute(filter_cols_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, ColRadix<RR>::kSmem));
                      PLAN_TRY(cudaFuncSetAttribute(dk_cols_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, ColRadix<RR>::kSmem)););
  }
  // streams / events of bffc_fwd_host (copy-in, compute, copy-out; per-slot events)
  for (auto& st : p->hs) PLAN_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (auto& ev : p->hev) PLAN_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
#undef PLAN_TRY
  *out = p;
  return BFFC_OK;
}

int bffc_fft_size(const bffc_plan* p) { return p ? p->NE : 0; }

int bffc_length_multiple(const bffc_plan* p) {
  if (!p) return 0;
  if (p->nlev == 0) return 64;                       // fused kernel: TMA tiles of 64 columns
  if (p->lev[0].tc) return p->N / 128;               // tcgen05 outer stage: whole rows of the [128][N/128] view
  return 8;                                          // CUDA-core outer stage: 16-byte vectors
}

int bffc_plan_destroy(bffc_plan* p) {
  if (!p) return BFFC_OK;
  cudaFree(p->dftC);
  cudaFree(p->dftS);
  cudaFree(p->gtiles);
  cudaFree(p->tw8192);
  cudaFree(p->tw512);
  cudaFree(p->tw_lo);
  cudaFree(p->tw_hi);
  for (auto& st : p->hs) if (st) cudaStreamDestroy(st);
  for (auto& ev : p->hev) if (ev) cudaEventDestroy(ev);
  delete p;
  return BFFC_OK;
}

int bffc_kf_pack(const bffc_plan* p, const void* kf_natural, void* kf_engine, int H, int conj, void* stream) {
  if (!p || !kf_natural || !kf_engine || H <= 0) return fail(BFFC_ERR_INVALID, "bffc_kf_pack: bad argument");
  if (p->nlev >= 1 && p->lev[0].R >= 32) {
    dim3 grid(kInner / 2 / 32, (p->lev[0].R / 32) * (p->nlev == 2 ? p->lev[1].R : 1), H);
    FMT_SWITCH(p->dtype, (kf_pack_tiled_kernel<false, F><<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(kf_natural), static_cast<uint2*>(kf_engine), p->NE, p->lev[0].R,
        p->nlev == 2 ? p->lev[1].R : 1, p->dtype == BFFC_DTYPE_BF16 ? 1.0f / float(p->NE) : 1.0f, conj)););
    CUDA_TRY(cudaGetLastError());
    return BFFC_OK;
  }
  dim3 grid((p->NE / 4 + 255) / 256 > 32 ? 32 : (p->NE / 4 + 255) / 256, H);
  FMT_SWITCH(p->dtype, (kf_pack_kernel<false, F><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float2*>(kf_natural), static_cast<uint4*>(kf_engine), p->NE, p->nlev >= 1 ? p->lev[0].R : 1,
      p->nlev == 2 ? p->lev[1].R : 1, p->dtype == BFFC_DTYPE_BF16 ? 1.0f / float(p->N) : 1.0f, conj,
      p->N < kInner ? p->N / 64 : 128)););
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

int bffc_kf_pack_rfft(const bffc_plan* p, const void* kf_half, void* kf_engine, int H, int conj, void* stream) {
  if (!p || !kf_half || !kf_engine || H <= 0) return fail(BFFC_ERR_INVALID, "bffc_kf_pack_rfft: bad argument");
  if (p->nlev >= 1 && p->lev[0].R >= 32) {
    dim3 grid(kInner / 2 / 32, (p->lev[0].R / 32) * (p->nlev == 2 ? p->lev[1].R : 1), H);
    FMT_SWITCH(p->dtype, (kf_pack_tiled_kernel<true, F><<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(kf_half), static_cast<uint2*>(kf_engine), p->NE, p->lev[0].R,
        p->nlev == 2 ? p->lev[1].R : 1, p->dtype == BFFC_DTYPE_BF16 ? 1.0f / float(p->NE) : 1.0f, conj)););
    CUDA_TRY(cudaGetLastError());
    return BFFC_OK;
  }
  dim3 grid((p->NE / 4 + 255) / 256 > 32 ? 32 : (p->NE / 4 + 255) / 256, H);
  FMT_SWITCH(p->dtype, (kf_pack_kernel<true, F><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float2*>(kf_half), static_cast<uint4*>(kf_engine), p->NE, p->nlev >= 1 ? p->lev[0].R : 1,
      p->nlev == 2 ? p->lev[1].R : 1, p->dtype == BFFC_DTYPE_BF16 ? 1.0f / float(p->N) : 1.0f, conj,
      p->N < kInner ? p->N / 64 : 128)););
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

// ---------------------------------------------------------------------------------------------- filter-side FFTs
// composite sizes: T = (channels, R/2 + 1, 8192) fp32 complex between the column and the row launch; the host walks the
// channels in groups whose T fits the workspace (bffc_filter_workspace_bytes sizes it to stay in L2)
static size_t filter_pair_bytes(const bffc_plan* p) { return size_t(2) * (p->R / 2 + 1) * kInner * sizeof(float2); }

size_t bffc_filter_workspace_bytes(const bffc_plan* p, int H) {
  if (!p || H <= 0 || p->NE == kInner) return 0;
  const size_t pairs = size_t(H + 1) / 2, per = filter_pair_bytes(p);
  // one group when it fits 576 MB (C4: 128 channels x 65 rows x 64 KB = 520 MB): the two launches are latency / issue
  // bound rather than HBM bound, so fewer, larger launches beat L2-sized groups (profiles/r2_filter_fft.md)
  size_t group = (size_t(576) << 20) / per;
  if (group < 1) group = 1;
  return (pairs < group ? pairs : group) * per;
}

int bffc_kf_from_filter(const bffc_plan* p, const void* k, int Lk, void* kf_engine, int H, int conj, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (!p || !k || !kf_engine || H <= 0 || Lk <= 0) return fail(BFFC_ERR_INVALID, "bffc_kf_from_filter: bad argument");
  if (Lk > p->N) return fail(BFFC_ERR_INVALID, "bffc_kf_from_filter: Lk=%d exceeds seqlen %d", Lk, p->N);
  using namespace bffc::ffft;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float scale = p->dtype == BFFC_DTYPE_BF16 ? 1.0f / float(p->N) : 1.0f;
  g_launches = 0;
  if (p->NE == kInner) {
    FMT_SWITCH(p->dtype, (kf_from_filter_kernel<F><<<(H + 1) / 2, kThreads, kSmemBytes, st>>>(
        static_cast<const float*>(k), Lk, static_cast<uint4*>(kf_engine), H, scale, conj, p->tw8192, p->N < kInner ? p->N : kInner)););
    CUDA_TRY(cudaGetLastError());
    g_launches = 1;
    return BFFC_OK;
  }
  const size_t per = filter_pair_bytes(p);
  if (!workspace || workspace_bytes < per)
    return fail(BFFC_ERR_INVALID, "bffc_kf_from_filter: workspace %zu B < %zu B (one channel pair; see bffc_filter_workspace_bytes)", workspace_bytes, per);
  const int group = int(std::min<size_t>(workspace_bytes / per, size_t(H + 1) / 2)) * 2;   // channels per group
  const int R = p->R, R0 = p->lev[0].R, R1 = p->nlev >= 2 ? p->lev[1].R : 1;
  float2* T = static_cast<float2*>(workspace);
  for (int h0 = 0; h0 < H; h0 += group) {
    const int Hc = std::min(group, H - h0);
    const float* kc = static_cast<const float*>(k) + size_t(h0) * Lk;
    uint4* out = static_cast<uint4*>(kf_engine) + size_t(h0) * (p->NE / 4);
    COLS_SWITCH(R, (filter_cols_kernel<RR><<<dim3(kInner / ColRadix<RR>::kTC, (Hc + 1) / 2), kColThreads, ColRadix<RR>::kSmem, st>>>(
        kc, Lk, T, Hc, scale, p->tw512, p->tw_lo, p->tw_hi)););
    FMT_SWITCH(p->dtype, (filter_rows_kernel<F><<<dim3(R / 2 + 1, Hc), kThreads, kSmemBytes, st>>>(T, out, R, R0, R1, conj, p->tw8192)););
    g_launches += 2;
  }
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

int bffc_dk_from_dkf(const bffc_plan* p, const void* dkf_engine, void* dk, int Lk, int H, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (!p || !dkf_engine || !dk || H <= 0 || Lk <= 0) return fail(BFFC_ERR_INVALID, "bffc_dk_from_dkf: bad argument");
  if (Lk > p->N) return fail(BFFC_ERR_INVALID, "bffc_dk_from_dkf: Lk=%d exceeds seqlen %d", Lk, p->N);
  using namespace bffc::ffft;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // fp16: both spectra carry 1/sqrt(128) (fused kernel) and 1/sqrt(R) per outer level: undo the product
  const float fscale = p->dtype == BFFC_DTYPE_BF16 ? 1.0f : float(p->N < kInner ? p->N / 64 : 128) * float(p->R);
  g_launches = 0;
  if (p->NE == kInner) {
    dk_from_dkf_kernel<<<H, kThreads, kSmemBytes, st>>>(static_cast<const float2*>(dkf_engine), static_cast<float*>(dk), Lk,
                                                         fscale, p->N < kInner ? p->N : kInner, p->tw8192);
    CUDA_TRY(cudaGetLastError());
    g_launches = 1;
    return BFFC_OK;
  }
  const size_t per = filter_pair_bytes(p);
  if (!workspace || workspace_bytes < per)
    return fail(BFFC_ERR_INVALID, "bffc_dk_from_dkf: workspace %zu B < %zu B (one channel pair; see bffc_filter_workspace_bytes)", workspace_bytes, per);
  const int group = int(std::min<size_t>(workspace_bytes / per, size_t(H + 1) / 2)) * 2;
  const int R = p->R, R0 = p->lev[0].R, R1 = p->nlev >= 2 ? p->lev[1].R : 1;
  float2* T = static_cast<float2*>(workspace);
  for (int h0 = 0; h0 < H; h0 += group) {
    const int Hc = std::min(group, H - h0);
    const float2* in = static_cast<const float2*>(dkf_engine) + size_t(h0) * p->NE;
    float* out = static_cast<float*>(dk) + size_t(h0) * Lk;
    dk_rows_kernel<<<dim3(R / 2 + 1, Hc), kThreads, kSmemBytes, st>>>(in, T, R, R0, R1, p->tw8192, p->tw_lo, p->tw_hi);
    COLS_SWITCH(R, (dk_cols_kernel<RR><<<dim3(kInner / ColRadix<RR>::kTC, (Hc + 1) / 2), kColThreads, ColRadix<RR>::kSmem, st>>>(
        T, out, Lk, Hc, fscale / float(p->NE), p->tw512)););
    g_launches += 2;
  }
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

int bffc_dkf_unpack(const bffc_plan* p, const void* dkf_engine, void* dkf_natural, int H, void* stream) {
  if (!p || !dkf_engine || !dkf_natural || H <= 0) return fail(BFFC_ERR_INVALID, "bffc_dkf_unpack: bad argument");
  if (p->N < kInner) {
    bffc::r128::dkf_unpack_small_kernel<<<dim3(kInner / 256, H), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(dkf_engine), static_cast<float2*>(dkf_natural), p->N,
        p->dtype == BFFC_DTYPE_BF16 ? 1.0f : float(p->N / 64), 0);
    CUDA_TRY(cudaGetLastError());
    return BFFC_OK;
  }
  dim3 grid(64, H);
  bffc::r128::dkf_unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float2*>(dkf_engine), static_cast<float2*>(dkf_natural), p->NE,
      p->nlev >= 1 ? p->lev[0].R : 1, p->nlev >= 2 ? p->lev[1].R : 1,
      // fp16: both spectra carry 1/sqrt(128) (fused kernel) and 1/sqrt(R) per outer level: undo the product
      p->dtype == BFFC_DTYPE_BF16 ? 1.0f : 128.0f * float(p->R));
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

int bffc_dkf_unpack_half(const bffc_plan* p, const void* dkf_engine, void* dkf_half, int H, void* stream) {
  if (!p || !dkf_engine || !dkf_half || H <= 0) return fail(BFFC_ERR_INVALID, "bffc_dkf_unpack_half: bad argument");
  if (p->N < kInner) {
    bffc::r128::dkf_unpack_small_kernel<<<dim3(kInner / 2 / 256 + 1, H), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(dkf_engine), static_cast<float2*>(dkf_half), p->N,
        p->dtype == BFFC_DTYPE_BF16 ? 1.0f : float(p->N / 64), 1);
    CUDA_TRY(cudaGetLastError());
    return BFFC_OK;
  }
  const int R0 = p->nlev >= 1 ? p->lev[0].R : 1, R1 = p->nlev >= 2 ? p->lev[1].R : 1, R = R0 * R1;
  dim3 grid(R < 32 ? 1 : R / 32, 128 * 2, H);
  bffc::r128::dkf_unpack_half_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float2*>(dkf_engine), static_cast<float2*>(dkf_half), p->NE, R0, R1,
      p->dtype == BFFC_DTYPE_BF16 ? 1.0f : 128.0f * float(p->R));       // fp16: as bffc_dkf_unpack
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

}  // extern "C"

// Composite sizes keep the outer stage's output as two bf16 planes (real, imaginary) of pairs*H*N elements.
static size_t plane_bytes(const bffc_plan* p, int B, int H) { return size_t((B + 1) / 2) * H * p->N * 2; }
// One launch group of a composite size works on batch members [0, B) (tensor pointers are pre-offset to the first one)
// and channels [h0, h0 + H) of (.., Hs, L) tensors; its plane sets hold ceil(B/2) * H * N complex elements.
struct View { int B, H, Hs, h0; };

// Composite sizes can run chunk by chunk: `sets` plane sets of one chunk take at most ~kPlaneBudget bytes.  Channels
// first (the k_f rows of a channel are then shared by its batch pairs inside one chunk); a single channel that is too
// large is cut over batch pairs.  Measured (profiles/r2_chunking.md): L2-sized chunks (40 MB, planes resident in the
// 126 MB L2) LOSE — C3 0.72 -> 1.00 ms, C4 0.85 -> 1.23 ms, C5 1.16 -> 2.52 ms — because every chunk pays the launch
// gaps, prologues and tails of 3-5 small persistent launches.  The budget therefore only bounds the workspace (and with
// it the peak memory of a call): 4 GB of plane sets per chunk.
constexpr size_t kPlaneBudget = size_t(4) << 30;
static View chunk_view(const bffc_plan* p, int B, int H, int sets) {
  const size_t item = size_t(sets) * p->N * 4;                     // one (pair, channel) in all plane sets
  size_t items = kPlaneBudget / item;
  if (items < 1) items = 1;
  const int pairs = (B + 1) / 2;
  View v{B, H, H, 0};
  if (items >= size_t(pairs)) {
    const size_t hc = items / pairs;
    v.H = hc < size_t(H) ? int(hc) : H;
  } else {
    v.H = 1;
    v.B = 2 * int(items);
    if (v.B > B) v.B = B;
  }
  return v;
}


// u*pregate and dout*postgate for the dk_f kernel of a gated backward at seqlen <= 8192 (two (B,H,L) tensors)
static size_t gate_scratch_bytes(int B, int H, int L) { return 2 * ((size_t(B) * H * L * 2 + 255) & ~size_t(255)); }

extern "C" size_t bffc_workspace_bytes_ex(const bffc_plan* p, int B, int H, int L, int gated, int backward) {
  if (!p) return 0;
  if (p->nlev == 0) return (gated && backward) ? gate_scratch_bytes(B, H, L) : 0;
  // plane sets (real + imaginary plane each) of ONE chunk: forward nlev sets; backward nlev + 1 (transformed u and dout)
  const View vf = chunk_view(p, B, H, p->nlev);
  size_t need = size_t(2 * p->nlev) * plane_bytes(p, vf.B, vf.H);
  if (backward) {      // the du / dpostgate passes run as forward chunks, the dk_f part with one more set per chunk
    const View vb = chunk_view(p, B, H, p->nlev + 1);
    const size_t nb = size_t(2 * (p->nlev + 1)) * plane_bytes(p, vb.B, vb.H);
    if (nb > need) need = nb;
  }
  return need;
}

extern "C" size_t bffc_workspace_bytes(const bffc_plan* p, int B, int H, int L) {
  return bffc_workspace_bytes_ex(p, B, H, L, 1, 1);     // enough for every call with these shapes
}

static CUtensorMapDataType map_dtype(const bffc_plan* p) {
  return p->dtype == BFFC_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

static int make_map(const bffc_plan* p, CUtensorMap* map, const void* base, int rows, int L, int box_rows = 128) {
  // (rows, L) bf16 viewed as [row][L/64][64]; box = one (128 x 64) tile, 128B swizzle;
  // tile rows >= L/64 are out of bounds: zero-filled on load (implicit padding), dropped on store.
  cuuint64_t dims[3] = {64, cuuint64_t(L / 64), cuuint64_t(rows)};
  cuuint64_t strides[2] = {128, cuuint64_t(L) * 2};
  cuuint32_t box[3] = {64, cuuint32_t(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, map_dtype(p), 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BFFC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(r));
  return 0;
}

static void fill_params(const bffc_plan* p, bffc::FwdParams& prm, const void* kf) {
  prm.kf = static_cast<const uint32_t*>(kf);
  prm.dftC = p->dftC;
  prm.dftS = p->dftS;
  prm.gtiles = p->gtiles;
  // Normalisation.  bf16: the whole 1/N is folded into k_f by the pack kernel (as the reference folds it into a
  // twiddle table, conv.py:146).  fp16 cannot hold k_f/N (underflow): every DFT stage is scaled by 1/sqrt(radix)
  // instead so intermediates stay near the input level: 1/sqrt(stage-1 radix) in the twiddle tables (passes 1 and 5),
  // 1/64 with the (unscaled) k_f in pass 3, 1/sqrt(R) per direction in the outer stages.
  prm.kf_scale = 1.0f / 64.0f;
  const int rblk = p->N < kInner ? p->N / 64 : 128;            // stage-1 radix
  prm.tw_scale = p->dtype == BFFC_DTYPE_BF16 ? 1.0f : 1.0f / sqrtf(float(rblk));
  prm.tw_n = rblk * 64;
  prm.tw_mask = rblk - 1;
  prm.pregate = nullptr;
  prm.postgate = nullptr;
  prm.postgate2 = nullptr;
  prm.y2 = nullptr;
  prm.xg_out = nullptr;
  prm.kf_conj_mask = 0;
  prm.trace = nullptr;
}

#ifdef BFFC_BRINGUP
// bring-up builds only (nvcc -DBFFC_BRINGUP): BFFC_TRACE=<file> makes every ungated fused launch record the phase
// timeline of CTA 0 (tools/trace_fwd3.py).  Not compiled into the product library.
static long long* g_trace = nullptr;
static long long* trace_buffer() {
  static bool init = false;
  if (!init) {
    init = true;
    if (getenv("BFFC_TRACE")) cudaMalloc(&g_trace, 3 * 2 * 64 * 16 * sizeof(long long));
  }
  if (g_trace) cudaMemset(g_trace, 0, 3 * 2 * 64 * 16 * sizeof(long long));
  return g_trace;
}
static void trace_dump(cudaStream_t st) {
  std::vector<long long> h(3 * 2 * 64 * 16);
  cudaStreamSynchronize(st);
  cudaMemcpy(h.data(), g_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  if (FILE* f = fopen(getenv("BFFC_TRACE"), "wb")) { fwrite(h.data(), sizeof(long long), h.size(), f); fclose(f); }
}
#endif

// Segment geometry of the input tiles: S = 8192/N batch members per 8192-point unit for the small sizes, else 1.
struct SegGeom { int S, seg_rows, groups, kmask; };
static SegGeom seg_geom(const bffc_plan* p, int B, int L) {
  SegGeom g;
  g.S = p->N < kInner ? kInner / p->N : 1;
  g.seg_rows = 128 / g.S;
  g.groups = (B + 2 * g.S - 1) / (2 * g.S);
  g.kmask = 0;
  const int used = (L + 63) / 64;                 // non-zero 64-element rows per segment
  for (int r = 0; r < 128; ++r)
    if (r % g.seg_rows < used) g.kmask |= 1 << (r / 16);
  return g;
}

// optional extras of one pass of the forward path
struct PassOpts {
  int conj = 0;                     // 1: conjugate k_f inside the kernel's pointwise multiply (else kf is pre-conjugated)
  const void* postgate2 = nullptr;  // second gated output y2 = postgate2 * conv(...) from the same pass
  void* y2 = nullptr;
  void* xg_out = nullptr;           // seqlen <= 8192, gated: the pass also stores its gated input u * pregate here
};

// fused 8192-point kernel on (B, H, L) real sequences (seqlen <= 8192)
static int launch_fused(const bffc_plan* p, const void* u, const void* kf, const void* pregate, const void* postgate,
                        void* y, int B, int H, int L, cudaStream_t st,
                        const PassOpts& po = PassOpts()) {
  if (L % 64 != 0) return fail(BFFC_ERR_UNSUPPORTED, "L=%d must be a multiple of 64 for seqlen <= 8192 in this build", L);
  const SegGeom sg = seg_geom(p, B, L);
  CUtensorMap tm_u, tm_y, tm_g;
  if (int rc = make_map(p, &tm_u, u, B * H, L, sg.seg_rows)) return rc;
  if (int rc = make_map(p, &tm_y, y, B * H, L, sg.seg_rows)) return rc;
  if (int rc = make_map(p, &tm_g, pregate ? pregate : u, B * H, L, sg.seg_rows)) return rc;
  bffc::FwdParams prm;
  fill_params(p, prm, kf);
  prm.pregate = static_cast<const uint32_t*>(pregate);
  prm.postgate = static_cast<const uint32_t*>(postgate);
  prm.postgate2 = static_cast<const uint32_t*>(po.postgate2);
  prm.y2 = static_cast<uint32_t*>(po.y2);
  prm.kf_conj_mask = po.conj ? 0x80008000u : 0u;
  prm.xg_out = pregate ? po.xg_out : nullptr;
  prm.B = B; prm.H = H; prm.L = L;
  prm.pairs = sg.groups;
  prm.kmask = sg.kmask;
  prm.nseg = sg.S;
  prm.seg_bytes = sg.seg_rows * 128;
  prm.units = H * prm.pairs;
  using namespace bffc::r128;
  const bool gated = pregate || postgate || prm.y2;
  // gate tiles travel by TMA like the inputs: pregate with the (segmented) geometry of u, output gates with that of y
  GateMaps gm{tm_g, tm_u, tm_u, tm_u, tm_u};
  if (prm.xg_out) { if (int rc = make_map(p, &gm.xg, prm.xg_out, B * H, L, sg.seg_rows)) return rc; }
  if (postgate) { if (int rc = make_map(p, &gm.post, postgate, B * H, L, sg.seg_rows)) return rc; }
  if (prm.y2) {
    if (int rc = make_map(p, &gm.post2, prm.postgate2, B * H, L, sg.seg_rows)) return rc;
    if (int rc = make_map(p, &gm.y2, prm.y2, B * H, L, sg.seg_rows)) return rc;
  }
  int g3 = (prm.units + kPipes3 - 1) / kPipes3;
  if (g3 > p->num_sms) g3 = p->num_sms;
  FMT_SWITCH(p->dtype,
    if (gated)
      fwd3_kernel<false, true, F><<<g3, kThreads3, kSmemTotal3, st>>>(tm_u, tm_y, tm_g, gm, prm);
    else {
#ifdef BFFC_BRINGUP
      prm.trace = trace_buffer();
#endif
      fwd3_kernel<false, false, F><<<g3, kThreads3, kSmemTotal3, st>>>(tm_u, tm_y, tm_g, gm, prm);
#ifdef BFFC_BRINGUP
      if (prm.trace) trace_dump(st);
#endif
    }
  );
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

// fused kernel on complex rows held in two planes (in place); rows = pairs * kf_rows
static int launch_planes(const bffc_plan* p, void* pre, void* pim, const void* kf, int pairs, int kf_rows, cudaStream_t st,
                         int conj = 0) {
  CUtensorMap tm_r, tm_i;
  if (int rc = make_map(p, &tm_r, pre, pairs * kf_rows, kInner)) return rc;
  if (int rc = make_map(p, &tm_i, pim, pairs * kf_rows, kInner)) return rc;
  bffc::FwdParams prm;
  fill_params(p, prm, kf);
  prm.B = 2 * pairs; prm.H = kf_rows; prm.L = kInner;
  prm.pairs = pairs;
  prm.kmask = 0xff; prm.nseg = 1; prm.seg_bytes = 16384;
  prm.units = kf_rows * pairs;
  prm.kf_conj_mask = conj ? 0x80008000u : 0u;
  using namespace bffc::r128;
  int g3 = (prm.units + kPipes3 - 1) / kPipes3;
  if (g3 > p->num_sms) g3 = p->num_sms;
  const GateMaps gm{tm_r, tm_r, tm_r, tm_r, tm_r};      // unused in this mode
  FMT_SWITCH(p->dtype, (fwd3_kernel<true, false, F><<<g3, kThreads3, kSmemTotal3, st>>>(tm_r, tm_r, tm_i, gm, prm)););
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

static int make_map4(const bffc_plan* p, CUtensorMap* map, const void* base, int chunks, int rows, int seqs,
                     size_t row_stride_bytes, size_t seq_stride_bytes) {
  // [seq][row][chunk][64] bf16 view of a strided matrix; box = 128 rows x 64 columns of one chunk, 128B swizzle.
  cuuint64_t dims[4] = {64, cuuint64_t(chunks), cuuint64_t(rows), cuuint64_t(seqs)};
  cuuint64_t strides[3] = {128, cuuint64_t(row_stride_bytes), cuuint64_t(seq_stride_bytes)};
  cuuint32_t box[4] = {64, 1, 128, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(map, map_dtype(p), 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BFFC_ERR_CUDA, "cuTensorMapEncodeTiled (4d) failed (%d)", int(r));
  return 0;
}

struct PlaneSet { uint8_t* re; uint8_t* im; };

static void fill_step(bffc::outer::OuterParams& op, double n_level) {
  for (int t = 0; t < 8; ++t) {
    const double ang = -2.0 * 3.14159265358979323846 * t / n_level;
    op.step[t] = make_float2(float(cos(ang)), float(sin(ang)));
  }
}

template <int R, int F>
static void launch_cc(const bffc_plan* p, bool inverse, bool gated, bool planes, const bffc::outer::OuterParams& op_in,
                      int rows, cudaStream_t st) {
  using namespace bffc::outer;
  OuterParams op = op_in;
  // read-ahead distance = the number of resident blocks (one residency period ahead; measured optimum at C3:
  // 0 -> 0.827 ms, 296 -> 0.690, 592 -> 0.693, 1184 -> 0.725, 2368 -> 0.985 ms for the three kernels)
  op.lookahead = p->num_sms * (R <= 4 ? 4 : 2);
  const int cb = op.M / (kVec * 128);
  if (planes) {
    dim3 grid(rows, cb, 1);
    if (!inverse) fwd_kernel<R, false, true, F><<<grid, 128, 0, st>>>(op);
    else inv_kernel<R, false, true, F><<<grid, 128, 0, st>>>(op);
  } else {
    dim3 grid(cb, op.H, op.pairs);
    if (!inverse) {
      if (gated) fwd_kernel<R, true, false, F><<<grid, 128, 0, st>>>(op);
      else fwd_kernel<R, false, false, F><<<grid, 128, 0, st>>>(op);
    } else {
      if (gated) inv_kernel<R, true, false, F><<<grid, 128, 0, st>>>(op);
      else inv_kernel<R, false, false, F><<<grid, 128, 0, st>>>(op);
    }
  }
}
static int cc_stage(const bffc_plan* p, int R, bool inverse, bool gated, bool planes, const bffc::outer::OuterParams& op, int rows,
                    cudaStream_t st) {
  switch (R) {
    case 2: FMT_SWITCH(p->dtype, (launch_cc<2, F>(p, inverse, gated, planes, op, rows, st));); break;
    case 4: FMT_SWITCH(p->dtype, (launch_cc<4, F>(p, inverse, gated, planes, op, rows, st));); break;
    case 8: FMT_SWITCH(p->dtype, (launch_cc<8, F>(p, inverse, gated, planes, op, rows, st));); break;
    default: return fail(BFFC_ERR_UNSUPPORTED, "outer radix %d not supported", R);
  }
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

// tcgen05 radix-128 level 0: real endpoint x (u or y), gate g (pregate fwd / postgate inv), planes set A
static int tc_stage(const bffc_plan* p, bool inverse, const void* x, const void* gate, PlaneSet A, View v, int L,
                    cudaStream_t st, const void* gate2 = nullptr, void* x2 = nullptr) {
  const int B = v.B, H = v.H;
  const int M = p->N / 128, chunks = M / 64, pairs = (B + 1) / 2;
  if (L % M != 0) return fail(BFFC_ERR_UNSUPPORTED, "seqlen %d needs L to be a multiple of %d in this build (L=%d)", p->N, M, L);
  CUtensorMap tm_x, tm_pr, tm_pi, tm_g;
  if (int rc = make_map4(p, &tm_x, x, chunks, L / M, B * v.Hs, size_t(M) * 2, size_t(L) * 2)) return rc;
  if (int rc = make_map4(p, &tm_pr, A.re, chunks, 128, pairs * H, size_t(M) * 2, size_t(p->N) * 2)) return rc;
  if (int rc = make_map4(p, &tm_pi, A.im, chunks, 128, pairs * H, size_t(M) * 2, size_t(p->N) * 2)) return rc;
  if (int rc = make_map4(p, &tm_g, (!inverse && gate) ? gate : x, chunks, L / M, B * v.Hs, size_t(M) * 2, size_t(L) * 2)) return rc;
  bffc::OuterTcParams prm;
  prm.dftC = p->dftC; prm.dftS = p->dftS;
  prm.postgate = inverse ? static_cast<const uint32_t*>(gate) : nullptr;
  prm.postgate2 = inverse ? static_cast<const uint32_t*>(gate2) : nullptr;
  prm.y2 = inverse ? static_cast<uint32_t*>(x2) : nullptr;
  prm.has_pregate = (!inverse && gate) ? 1 : 0;
  prm.tw_scale = p->dtype == BFFC_DTYPE_BF16 ? 1.0f : 0.08838834764831845f;
  prm.B = B; prm.H = H; prm.L = L; prm.pairs = pairs;
  prm.Hs = v.Hs; prm.h0 = v.h0;
  prm.N = p->N; prm.M = M; prm.chunks = chunks;
  prm.ksteps = (L / M + 15) / 16;
  prm.units = pairs * H * chunks;
  int grid = (prm.units + 1) / 2;
  if (grid > p->num_sms) grid = p->num_sms;
  using namespace bffc::r128;
  FMT_SWITCH(p->dtype,
    if (!inverse)
      outer_tc_kernel<false, F><<<grid, kThreads, prm.has_pregate ? kSmemOuterGated : kSmemOuter, st>>>(tm_x, tm_pr, tm_pi, tm_g, prm);
    else
      outer_tc_kernel<true, F><<<grid, kThreads, kSmemOuter, st>>>(tm_x, tm_pr, tm_pi, tm_g, prm);
  );
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

static PlaneSet plane_set(const bffc_plan* p, void* ws, int idx, int B, int H) {
  uint8_t* b = static_cast<uint8_t*>(ws) + size_t(2 * idx) * plane_bytes(p, B, H);
  return PlaneSet{b, b + plane_bytes(p, B, H)};
}

// all outer levels, forward: real (B,H,L) x (* pregate) -> complex 8192-point rows.  Level 0 writes set `s0`,
// level 1 (if any) reads `s0` and writes `s1`; returns the set holding the rows.
static int transform_fwd(const bffc_plan* p, const void* x, const void* pregate, View v, int L, PlaneSet s0,
                         PlaneSet s1, PlaneSet* out, cudaStream_t st, int* launches) {
  const int B = v.B, H = v.H;
  const int pairs = (B + 1) / 2;
  const bffc_level l0 = p->lev[0];
  if (l0.tc) {
    if (int rc = tc_stage(p, false, x, pregate, s0, v, L, st)) return rc;
  } else {
    bffc::outer::OuterParams op{};
    op.u = static_cast<const uint4*>(x);
    op.pregate = static_cast<const uint4*>(pregate);
    op.pre = reinterpret_cast<uint4*>(s0.re); op.pim = reinterpret_cast<uint4*>(s0.im);
    op.B = B; op.H = H; op.L = L; op.pairs = pairs; op.M = p->N / l0.R;
    op.Hs = v.Hs; op.h0 = v.h0;
    op.scale = p->dtype == BFFC_DTYPE_BF16 ? 1.0f : 1.0f / sqrtf(float(l0.R));
    fill_step(op, double(p->N));
    if (int rc = cc_stage(p, l0.R, false, pregate != nullptr, false, op, 0, st)) return rc;
  }
  *launches += 1;
  *out = s0;
  if (p->nlev == 2) {
    const bffc_level l1 = p->lev[1];
    bffc::outer::OuterParams op{};
    op.xre = reinterpret_cast<uint4*>(s0.re); op.xim = reinterpret_cast<uint4*>(s0.im);
    op.pre = reinterpret_cast<uint4*>(s1.re); op.pim = reinterpret_cast<uint4*>(s1.im);
    op.B = B; op.H = H; op.L = L; op.pairs = pairs; op.M = p->N / (l0.R * l1.R);
    op.scale = p->dtype == BFFC_DTYPE_BF16 ? 1.0f : 1.0f / sqrtf(float(l1.R));
    fill_step(op, double(p->N) / l0.R);
    if (int rc = cc_stage(p, l1.R, false, false, true, op, pairs * H * l0.R, st)) return rc;
    *launches += 1;
    *out = s1;
  }
  return BFFC_OK;
}

// all outer levels, inverse: rows in `rows` (set s1 if two levels, else s0) -> real y (* postgate)
static int transform_inv(const bffc_plan* p, void* y, const void* postgate, View v, int L, PlaneSet s0, PlaneSet s1,
                         cudaStream_t st, int* launches, const void* postgate2 = nullptr, void* y2 = nullptr) {
  const int B = v.B, H = v.H;
  const int pairs = (B + 1) / 2;
  const bffc_level l0 = p->lev[0];
  if (p->nlev == 2) {
    const bffc_level l1 = p->lev[1];
    bffc::outer::OuterParams op{};
    op.xre = reinterpret_cast<uint4*>(s0.re); op.xim = reinterpret_cast<uint4*>(s0.im);
    op.pre = reinterpret_cast<uint4*>(s1.re); op.pim = reinterpret_cast<uint4*>(s1.im);
    op.B = B; op.H = H; op.L = L; op.pairs = pairs; op.M = p->N / (l0.R * l1.R);
    op.scale = p->dtype == BFFC_DTYPE_BF16 ? 1.0f : 1.0f / sqrtf(float(l1.R));
    fill_step(op, double(p->N) / l0.R);
    if (int rc = cc_stage(p, l1.R, true, false, true, op, pairs * H * l0.R, st)) return rc;
    *launches += 1;
  }
  if (l0.tc) {
    if (int rc = tc_stage(p, true, y, postgate, s0, v, L, st, postgate2, y2)) return rc;
  } else {
    bffc::outer::OuterParams op{};
    op.y = static_cast<uint4*>(y);
    op.postgate = static_cast<const uint4*>(postgate);
    op.postgate2 = static_cast<const uint4*>(postgate2);
    op.y2 = static_cast<uint4*>(y2);
    op.pre = reinterpret_cast<uint4*>(s0.re); op.pim = reinterpret_cast<uint4*>(s0.im);
    op.B = B; op.H = H; op.L = L; op.pairs = pairs; op.M = p->N / l0.R;
    op.Hs = v.Hs; op.h0 = v.h0;
    op.scale = p->dtype == BFFC_DTYPE_BF16 ? 1.0f : 1.0f / sqrtf(float(l0.R));
    fill_step(op, double(p->N));
    if (int rc = cc_stage(p, l0.R, true, postgate != nullptr, false, op, 0, st)) return rc;
  }
  *launches += 1;
  return BFFC_OK;
}

static int check_common(const bffc_plan* p, int B, int H, int L, const void* a, const void* b, const void* c) {
  if (!p) return fail(BFFC_ERR_INVALID, "null plan");
  // Tensor maps are encoded through the driver API, which needs a current context in the CALLING thread.  A thread that
  // has made no runtime call yet (PyTorch's autograd worker entering bffc_bwd) has none: this runtime no-op binds the
  // device's primary context (cuTensorMapEncodeTiled otherwise fails with CUDA_ERROR_INVALID_CONTEXT).
  CUDA_TRY(cudaFree(nullptr));
  if (B <= 0 || H <= 0 || L <= 0 || L > p->N) return fail(BFFC_ERR_INVALID, "bad shape B=%d H=%d L=%d (seqlen %d)", B, H, L, p->N);
  if (L % 8 != 0) return fail(BFFC_ERR_UNSUPPORTED, "L=%d must be a multiple of 8 in this build", L);
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15)
    return fail(BFFC_ERR_INVALID, "device pointers must be 16-byte aligned");
  return 0;
}

// y = postgate * conv(u * pregate, k) for any supported size.  `ws`: workspace (plane sets 0 and 1) for composite sizes.
static int conv_forward(const bffc_plan* p, const void* u, const void* kf, const void* pregate, const void* postgate,
                        void* y, int B, int H, int L, void* ws, cudaStream_t st, int* launches,
                        const PassOpts& po = PassOpts()) {
  if (p->nlev == 0) {
    *launches += 1;
    return launch_fused(p, u, kf, pregate, postgate, y, B, H, L, st, po);
  }
  // composite sizes, chunk by chunk (see chunk_view): outer stage(s) -> inner kernel in place -> inverse outer stage(s)
  const View c = chunk_view(p, B, H, p->nlev);
  const size_t bstride = size_t(H) * L * 2;                       // bytes of one batch member of the (B, H, L) tensors
  auto at = [&](const void* t, int b0) { return t ? static_cast<const uint8_t*>(t) + size_t(b0) * bstride : nullptr; };
  for (int b0 = 0; b0 < B; b0 += c.B)
    for (int h0 = 0; h0 < H; h0 += c.H) {
      const View v{B - b0 < c.B ? B - b0 : c.B, H - h0 < c.H ? H - h0 : c.H, H, h0};
      const int pairs = (v.B + 1) / 2;
      PlaneSet s0 = plane_set(p, ws, 0, c.B, c.H), s1 = plane_set(p, ws, p->nlev == 2 ? 1 : 0, c.B, c.H), rows;
      if (int rc = transform_fwd(p, at(u, b0), at(pregate, b0), v, L, s0, s1, &rows, st, launches)) return rc;
      const uint8_t* kfc = static_cast<const uint8_t*>(kf) + size_t(h0) * p->NE * 4;
      if (int rc = launch_planes(p, rows.re, rows.im, kfc, pairs, v.H * p->R, st, po.conj)) return rc;
      *launches += 1;
      if (int rc = transform_inv(p, const_cast<uint8_t*>(at(y, b0)), at(postgate, b0), v, L, s0, s1, st, launches,
                                 at(po.postgate2, b0), const_cast<uint8_t*>(at(po.y2, b0)))) return rc;
    }
  return BFFC_OK;
}

extern "C" {

int bffc_fwd(const bffc_plan* p, const void* u, const void* kf, const void* pregate, const void* postgate, void* y,
             int B, int H, int L, void* workspace, size_t workspace_bytes, void* stream) {
  if ((pregate == nullptr) != (postgate == nullptr))
    return fail(BFFC_ERR_INVALID, "bffc_fwd: pregate and postgate must both be given or both be null");
  if (!u || !kf || !y) return fail(BFFC_ERR_INVALID, "bffc_fwd: null pointer");
  if (int rc = check_common(p, B, H, L, u, y, kf)) return rc;
  if ((reinterpret_cast<uintptr_t>(pregate) | reinterpret_cast<uintptr_t>(postgate) | reinterpret_cast<uintptr_t>(workspace)) & 15)
    return fail(BFFC_ERR_INVALID, "bffc_fwd: gates / workspace must be 16-byte aligned");
  {
    const size_t need = bffc_workspace_bytes_ex(p, B, H, L, pregate != nullptr, 0);
    if (need && (!workspace || workspace_bytes < need))
      return fail(BFFC_ERR_INVALID, "bffc_fwd: workspace of %zu bytes required", need);
  }
  int launches = 0;
  int rc = conv_forward(p, u, kf, pregate, postgate, y, B, H, L, workspace, static_cast<cudaStream_t>(stream), &launches);
  g_launches = launches;
  return rc;
}

int bffc_bwd(const bffc_plan* p, const void* dout, const void* u, const void* kf, const void* kf_conj,
             const void* pregate, const void* postgate, void* du, void* dkf, void* dpregate, void* dpostgate, int B, int H,
             int L, void* workspace, size_t workspace_bytes, void* stream) {
  if ((pregate == nullptr) != (postgate == nullptr))
    return fail(BFFC_ERR_INVALID, "bffc_bwd: pregate and postgate must both be given or both be null");
  const bool gated = pregate != nullptr;
  if (!dout || !u || (!kf && !kf_conj) || !du || !dkf) return fail(BFFC_ERR_INVALID, "bffc_bwd: null pointer");
  if (gated && (!kf || !dpregate || !dpostgate)) return fail(BFFC_ERR_INVALID, "bffc_bwd: gated backward needs kf, dpregate, dpostgate");
  if ((reinterpret_cast<uintptr_t>(pregate) | reinterpret_cast<uintptr_t>(postgate) | reinterpret_cast<uintptr_t>(dpregate) |
       reinterpret_cast<uintptr_t>(dpostgate) | reinterpret_cast<uintptr_t>(kf)) & 15)
    return fail(BFFC_ERR_INVALID, "bffc_bwd: gate pointers must be 16-byte aligned");
  if (int rc = check_common(p, B, H, L, u, du, dout)) return rc;
  if ((reinterpret_cast<uintptr_t>(dkf) | reinterpret_cast<uintptr_t>(kf_conj) | reinterpret_cast<uintptr_t>(workspace)) & 15)
    return fail(BFFC_ERR_INVALID, "bffc_bwd: dkf / kf / workspace must be 16-byte aligned");
  {
    const size_t need = bffc_workspace_bytes_ex(p, B, H, L, gated, 1);
    if (need && (!workspace || workspace_bytes < need))
      return fail(BFFC_ERR_INVALID, "bffc_bwd: workspace of %zu bytes required", need);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int launches = 0;
  // du = corr(dout, k) = circular conv with conj(k_f): the forward path on dout, conjugating k_f in the kernel's
  // pointwise multiply (a pre-conjugated kf_engine_conj is still accepted)
  // (reference: kernels_bf16/monarch_cuda_32_16_16_bwd_kernel_bf16.h:740-815)
  PassOpts dx;
  dx.conj = kf_conj ? 0 : 1;
  const void* kfc = kf_conj ? kf_conj : kf;
  uint8_t *gate_x = nullptr, *gate_d = nullptr;
  if (p->nlev > 0) {
    // composite sizes: all passes share the transformed rows, chunk by chunk, below
  } else if (!gated) {
    if (int rc = conv_forward(p, dout, kfc, nullptr, nullptr, du, B, H, L, workspace, st, &launches, dx)) return rc;
  } else {
    // y = q * conv(u*p, k)  (conv.py:3856-3939; kernels_bf16/..._bwd_kernel_bf16.h:836-906; host recompute
    // monarch_cuda_interface_bwd_bf16.cu:798-808).  With dx = corr(dout*q, k):
    //   dpostgate = dout * conv(u*p, k)                        — one pass of the forward path
    //   du = p * dx  and  dpregate = u * dx                    — ONE more pass with two gated outputs
    // the dk_f kernel below needs u*p and dout*q — the gated inputs of these two passes, which store them into the
    // tail of the workspace on the way
    PassOpts p1;
    gate_x = static_cast<uint8_t*>(workspace);
    gate_d = gate_x + gate_scratch_bytes(B, H, L) / 2;
    p1.xg_out = gate_x;
    dx.xg_out = gate_d;
    if (int rc = conv_forward(p, u, kf, pregate, dout, dpostgate, B, H, L, workspace, st, &launches, p1)) return rc;
    dx.postgate2 = u;
    dx.y2 = dpregate;
    if (int rc = conv_forward(p, dout, kfc, postgate, pregate, du, B, H, L, workspace, st, &launches, dx)) return rc;
  }
  // dk_f = sum_b FFT(dout*q) * conj(FFT(u*p)), reduced with fp32 atomics into the zeroed gradient
  CUDA_TRY(cudaMemsetAsync(dkf, 0, size_t(H) * p->NE * sizeof(float2), st));
  const int pairs = (B + 1) / 2;
  bffc::DkfParams prm;
  prm.dftC = p->dftC;
  prm.dftS = p->dftS;
  prm.gtiles = p->gtiles;
  prm.dkf = static_cast<float2*>(dkf);
  prm.pairs = pairs;
  {
    const int rblk = p->N < kInner ? p->N / 64 : 128;          // stage-1 radix, as fill_params
    prm.tw_scale = p->dtype == BFFC_DTYPE_BF16 ? 1.0f : 1.0f / sqrtf(float(rblk));
    prm.tw_n = rblk * 64;
    prm.tw_mask = rblk - 1;
  }
  using namespace bffc::r128;
  if (p->nlev == 0) {
    if (L % 64 != 0) return fail(BFFC_ERR_UNSUPPORTED, "L=%d must be a multiple of 64 for seqlen <= 8192 in this build", L);
    // gated loads (reference: ..._bwd_kernel_bf16.h:505-509,571-581): the products stored by the two passes above
    const void *xu = gated ? gate_x : u, *xd = gated ? gate_d : dout;
    CUtensorMap tm_u, tm_d;
    const SegGeom sg = seg_geom(p, B, L);
    if (int rc = make_map(p, &tm_u, xu, B * H, L, sg.seg_rows)) return rc;
    if (int rc = make_map(p, &tm_d, xd, B * H, L, sg.seg_rows)) return rc;
    prm.B = B; prm.H = H; prm.L = L;
    prm.pairs = sg.groups;
    prm.kmask = sg.kmask; prm.nseg = sg.S; prm.seg_bytes = sg.seg_rows * 128;
    const long long units = (long long)H * prm.pairs;
    const int grid = int(units < p->num_sms ? units : p->num_sms);
    FMT_SWITCH(p->dtype, (dkf3_kernel<false, F><<<grid, kThreadsDkf3, kSmemTotalDkf3, st>>>(tm_u, tm_d, tm_u, tm_d, prm)););
    CUDA_TRY(cudaGetLastError());
    launches += 1;
  } else {
    // chunk by chunk: the outer stages turn u (* pregate) and dout (* postgate) into complex 8192-point rows ONCE
    // (set U, set D; set 0 is the level-0 intermediate when nlev == 2).  The dk_f kernel reads both sets first; the
    // convolution passes then run on the same rows in place — rows of dout with conj k_f -> inverse outer stages -> du
    // (gated: x pregate, and x u -> dpregate from the same pass), gated: rows of u with k_f -> x dout -> dpostgate
    // (conv.py:3856-3939; kernels_bf16/..._bwd_kernel_bf16.h:836-906).  Batch chunks of a channel add into the same dk_f rows.
    const View c = chunk_view(p, B, H, p->nlev + 1);
    const size_t bstride = size_t(H) * L * 2;
    auto at = [&](const void* t, int b0) { return t ? static_cast<const uint8_t*>(t) + size_t(b0) * bstride : nullptr; };
    auto atw = [&](void* t, int b0) { return t ? static_cast<uint8_t*>(t) + size_t(b0) * bstride : nullptr; };
    for (int b0 = 0; b0 < B; b0 += c.B)
      for (int h0 = 0; h0 < H; h0 += c.H) {
        const View v{B - b0 < c.B ? B - b0 : c.B, H - h0 < c.H ? H - h0 : c.H, H, h0};
        const int vpairs = (v.B + 1) / 2;
        PlaneSet s0 = plane_set(p, workspace, 0, c.B, c.H);
        PlaneSet sU = p->nlev == 2 ? plane_set(p, workspace, 1, c.B, c.H) : s0;
        PlaneSet sD = plane_set(p, workspace, p->nlev == 2 ? 2 : 1, c.B, c.H);
        PlaneSet ru, rd;
        if (int rc = transform_fwd(p, at(u, b0), at(pregate, b0), v, L, s0, sU, &ru, st, &launches)) return rc;
        if (int rc = transform_fwd(p, at(dout, b0), at(postgate, b0), v, L, p->nlev == 2 ? s0 : sD, sD, &rd, st, &launches)) return rc;
        const int rows = v.H * p->R;
        CUtensorMap tur, tui, tdr, tdi;
        if (int rc = make_map(p, &tur, ru.re, vpairs * rows, kInner)) return rc;
        if (int rc = make_map(p, &tui, ru.im, vpairs * rows, kInner)) return rc;
        if (int rc = make_map(p, &tdr, rd.re, vpairs * rows, kInner)) return rc;
        if (int rc = make_map(p, &tdi, rd.im, vpairs * rows, kInner)) return rc;
        prm.B = 2 * vpairs; prm.H = rows; prm.L = kInner;
        prm.pairs = vpairs;
        prm.dkf = static_cast<float2*>(dkf) + size_t(h0) * p->NE;
        prm.kmask = 0xff; prm.nseg = 1; prm.seg_bytes = 16384;
        const long long units = (long long)rows * vpairs;
        const int grid = int(units < p->num_sms ? units : p->num_sms);
        FMT_SWITCH(p->dtype, (dkf3_kernel<true, F><<<grid, kThreadsDkf3, kSmemTotalDkf3, st>>>(tur, tdr, tui, tdi, prm)););
        CUDA_TRY(cudaGetLastError());
        launches += 1;
        const size_t kf_off = size_t(h0) * p->NE * 4;
        if (gated) {
          if (int rc = launch_planes(p, ru.re, ru.im, static_cast<const uint8_t*>(kf) + kf_off, vpairs, rows, st, 0)) return rc;
          launches += 1;
          if (int rc = transform_inv(p, atw(dpostgate, b0), at(dout, b0), v, L, p->nlev == 2 ? s0 : sU, sU, st, &launches)) return rc;
        }
        if (int rc = launch_planes(p, rd.re, rd.im, static_cast<const uint8_t*>(kfc) + kf_off, vpairs, rows, st, dx.conj)) return rc;
        launches += 1;
        if (int rc = transform_inv(p, atw(du, b0), at(pregate, b0), v, L, p->nlev == 2 ? s0 : sD, sD, st, &launches,
                                   gated ? at(u, b0) : nullptr, gated ? atw(dpregate, b0) : nullptr)) return rc;
      }
  }
  g_launches = launches;
  return BFFC_OK;
}

// ---------------------------------------------------------------------------------------------- host streaming
// Chunk geometry of the host pipeline: bc batch members x hc channels per chunk, ~12 MB per chunk tensor — a copy of
// ~0.25 ms (PCIe 5 x16) still runs at link speed, and the fill / drain of the three-stage pipeline (one chunk copy-in
// before, one copy-out after the overlapped part) stays small.  bc is even so that batch pairs stay together; wide
// rows (H*L*2 bytes > 6 MB) are split over channels instead and moved with pitched (2-D) copies.
struct HostChunk { int bc, hc; };
static HostChunk host_chunk(int B, int H, int L) {
  const size_t target = size_t(12) << 20, row = size_t(H) * L * 2;
  HostChunk g{2, H};
  if (2 * row > target) {
    const int nh = int((2 * row + target - 1) / target);
    g.hc = (H + nh - 1) / nh;
  } else {
    g.bc = int((target / row) & ~size_t(1));
  }
  if (g.bc > B) g.bc = B;
  return g;
}

int bffc_host_chunk_batch(const bffc_plan* p, int B, int H, int L) {
  if (!p || B <= 0 || H <= 0 || L <= 0) return 0;
  return host_chunk(B, H, L).bc;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

size_t bffc_host_workspace_bytes(const bffc_plan* p, int B, int H, int L, int gated) {
  if (!p || B <= 0 || H <= 0 || L <= 0) return 0;
  const HostChunk g = host_chunk(B, H, L);
  const size_t t = align256(size_t(g.bc) * g.hc * L * 2);
  return 2 * ((gated ? 4 : 2) * t + align256(bffc_workspace_bytes_ex(p, g.bc, g.hc, L, gated, 0)));
}

int bffc_fwd_host(const bffc_plan* p, const void* u_host, const void* kf, const void* pre_host, const void* post_host,
                  void* y_host, int B, int H, int L, void* dev_ws, size_t dev_ws_bytes, void* stream) {
  if ((pre_host == nullptr) != (post_host == nullptr))
    return fail(BFFC_ERR_INVALID, "bffc_fwd_host: pregate and postgate must both be given or both be null");
  if (!u_host || !kf || !y_host) return fail(BFFC_ERR_INVALID, "bffc_fwd_host: null pointer");
  if (int rc = check_common(p, B, H, L, kf, dev_ws, nullptr)) return rc;
  const bool gated = pre_host != nullptr;
  if (!dev_ws || dev_ws_bytes < bffc_host_workspace_bytes(p, B, H, L, gated))
    return fail(BFFC_ERR_INVALID, "bffc_fwd_host: device workspace of %zu bytes required", bffc_host_workspace_bytes(p, B, H, L, gated));
  cudaStream_t user = static_cast<cudaStream_t>(stream), s_in = p->hs[0], s_cmp = p->hs[1], s_out = p->hs[2];
  cudaEvent_t ev_start = p->hev[0];
  const cudaEvent_t* ev_in = &p->hev[1];    // [slot] chunk copied in
  const cudaEvent_t* ev_cmp = &p->hev[3];   // [slot] chunk convolved (input slot free again)
  const cudaEvent_t* ev_out = &p->hev[5];   // [slot] chunk copied out (output slot free again)
  const HostChunk g = host_chunk(B, H, L);
  const size_t t = align256(size_t(g.bc) * g.hc * L * 2);
  const size_t conv_ws = bffc_workspace_bytes_ex(p, g.bc, g.hc, L, gated, 0);
  const size_t slot_bytes = (gated ? 4 : 2) * t + align256(conv_ws);
  const size_t host_pitch = size_t(H) * L * 2;                   // one batch member of the host tensors
  const size_t kf_row = size_t(p->NE) * 4;                       // one channel of kf_engine
  // Error handling: once copies are in flight a failure must not return before they have stopped touching the caller's
  // host / device buffers, so the body runs in a lambda and every exit joins the three internal streams.
  int launches = 0, c = 0;
  auto body = [&]() -> int {
    CUDA_TRY(cudaEventRecord(ev_start, user));
    CUDA_TRY(cudaStreamWaitEvent(s_in, ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(s_out, ev_start, 0));
    for (int b0 = 0; b0 < B; b0 += g.bc)
      for (int h0 = 0; h0 < H; h0 += g.hc, ++c) {
        const int nb = B - b0 < g.bc ? B - b0 : g.bc, nh = H - h0 < g.hc ? H - h0 : g.hc, slot = c & 1;
        uint8_t* base = static_cast<uint8_t*>(dev_ws) + slot * slot_bytes;
        uint8_t *d_u = base, *d_y = base + t, *d_p = gated ? base + 2 * t : nullptr, *d_q = gated ? base + 3 * t : nullptr;
        uint8_t* d_ws = base + (gated ? 4 : 2) * t;
        // chunk = rows b0..b0+nb of a (B, H*L) matrix, columns h0*L..(h0+nh)*L: pitched on the host, dense on the device
        const size_t off = size_t(b0) * host_pitch + size_t(h0) * L * 2, width = size_t(nh) * L * 2;
        if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(s_in, ev_cmp[slot], 0));
        CUDA_TRY(cudaMemcpy2DAsync(d_u, width, static_cast<const uint8_t*>(u_host) + off, host_pitch, width, nb, cudaMemcpyHostToDevice, s_in));
        if (gated) {
          CUDA_TRY(cudaMemcpy2DAsync(d_p, width, static_cast<const uint8_t*>(pre_host) + off, host_pitch, width, nb, cudaMemcpyHostToDevice, s_in));
          CUDA_TRY(cudaMemcpy2DAsync(d_q, width, static_cast<const uint8_t*>(post_host) + off, host_pitch, width, nb, cudaMemcpyHostToDevice, s_in));
        }
        CUDA_TRY(cudaEventRecord(ev_in[slot], s_in));
        CUDA_TRY(cudaStreamWaitEvent(s_cmp, ev_in[slot], 0));
        if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(s_cmp, ev_out[slot], 0));
        if (int rc = conv_forward(p, d_u, static_cast<const uint8_t*>(kf) + size_t(h0) * kf_row, d_p, d_q, d_y, nb, nh, L,
                                  conv_ws ? d_ws : nullptr, s_cmp, &launches)) return rc;
        CUDA_TRY(cudaEventRecord(ev_cmp[slot], s_cmp));
        CUDA_TRY(cudaStreamWaitEvent(s_out, ev_cmp[slot], 0));
        CUDA_TRY(cudaMemcpy2DAsync(static_cast<uint8_t*>(y_host) + off, host_pitch, d_y, width, width, nb, cudaMemcpyDeviceToHost, s_out));
        CUDA_TRY(cudaEventRecord(ev_out[slot], s_out));
      }
    // join: everything enqueued above is complete when the last copy-out is (s_out is in order and waited on each
    // chunk's compute, which waited on its copy-in)
    CUDA_TRY(cudaStreamWaitEvent(user, ev_out[(c - 1) & 1], 0));
    return BFFC_OK;
  };
  const int rc = body();
  if (rc != BFFC_OK) {      // keep the message of the failure; quiesce the internal streams before handing control back
    char msg[sizeof(g_err)];
    memcpy(msg, g_err, sizeof(msg));
    for (auto st : p->hs) cudaStreamSynchronize(st);
    cudaGetLastError();
    memcpy(g_err, msg, sizeof(msg));
    return rc;
  }
  g_launches = launches;
  return BFFC_OK;
}

}  // extern "C"

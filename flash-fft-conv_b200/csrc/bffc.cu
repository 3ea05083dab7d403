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
#include "fwd_r128.cuh"
#include "dkf_r128.cuh"
#include "outer_cuda.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
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

uint16_t f2bf(double x) {  // round-to-nearest-even float -> bf16 bits
  float f = static_cast<float>(x);
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t r = 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>((u + r) >> 16);
}

__global__ void kf_pack_kernel(const float* __restrict__ kf_nat, uint32_t* __restrict__ kf_eng,
                               const int* __restrict__ perm, int N, int pair_stride, float scale, int conj) {
  const int h = blockIdx.y;
  const float* src = kf_nat + size_t(h) * N * 2;      // interleaved (re, im) fp32
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < N; w += gridDim.x * blockDim.x) {
    const int pw = perm[w];
    const int part = pw & 1;
    const float sc = (part && conj) ? -scale : scale;
    const float a = src[pw] * sc;                      // element k      (re or im)
    const float b = src[pw + 2 * pair_stride] * sc;    // element k + pair_stride
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    kf_eng[size_t(h) * N + w] = *reinterpret_cast<uint32_t*>(&v);
  }
}

constexpr int kInner = 8192;   // the fused tcgen05 kernel's size

}  // namespace

struct bffc_plan {
  int N;
  int R;         // N = R * 8192; R > 1: outer radix-R stage on CUDA cores around the fused kernel
  int dtype;
  int device;
  __nv_bfloat16* dftC = nullptr;
  __nv_bfloat16* dftS = nullptr;
  uint8_t* gtiles = nullptr;
  int* perm = nullptr;  // engine word index -> 2 * natural frequency index + part
  int num_sms = 0;
};

extern "C" {

int bffc_abi_version(void) { return BFFC_ABI_VERSION; }
const char* bffc_last_error(void) { return g_err; }
int bffc_last_launch_count(void) { return g_launches; }

int bffc_supported(int seqlen, int dtype) {
  if (dtype != BFFC_DTYPE_BF16) return 0;
  return (seqlen == 8192 || seqlen == 16384 || seqlen == 32768 || seqlen == 65536) ? 1 : 0;
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
  p->N = seqlen;
  p->R = seqlen / kInner;
  p->dtype = dtype;
  CUDA_TRY(cudaGetDevice(&p->device));
  CUDA_TRY(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));

  const double PI = 3.14159265358979323846;
  // outer radix-128 DFT of the fused kernel, cos / sin planes (symmetric, K-major rows)
  std::vector<uint16_t> c(128 * 128), s(128 * 128);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 128; ++k) {
      const double ang = 2.0 * PI * double((m * k) & 127) / 128.0;
      c[m * 128 + k] = f2bf(cos(ang));
      s[m * 128 + k] = f2bf(sin(ang));
    }
  CUDA_TRY(cudaMalloc(&p->dftC, c.size() * 2));
  CUDA_TRY(cudaMalloc(&p->dftS, s.size() * 2));
  CUDA_TRY(cudaMemcpy(p->dftC, c.data(), c.size() * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(p->dftS, s.data(), s.size() * 2, cudaMemcpyHostToDevice));

  // DFT-64 planes for the row-local stage: G = exp(-2 pi i k n / 64) = Gr + i Gi.  Stored as the MN-major
  // B operand image: row k (K index) = 64 bf16 = 128 B, 16-byte chunk c of row k at chunk position c ^ (k & 7)
  // (the 128B swizzle TMA / UMMA use).
  std::vector<uint8_t> gt(4 * bffc::r128::kGTileBytes, 0);   // tiles: Gr, Gi, -Gi, Gr
  for (int k = 0; k < 64; ++k)
    for (int n = 0; n < 64; ++n) {
      const double ang = -2.0 * PI * double((k * n) & 63) / 64.0;
      const uint16_t gr = f2bf(cos(ang)), gi = f2bf(sin(ang)), ngi = f2bf(-sin(ang));
      const size_t off = size_t(k) * 128 + (size_t((n >> 3) ^ (k & 7)) << 4) + size_t(n & 7) * 2;
      const size_t T = bffc::r128::kGTileBytes;
      memcpy(gt.data() + off, &gr, 2);
      memcpy(gt.data() + T + off, &gi, 2);
      memcpy(gt.data() + 2 * T + off, &ngi, 2);
      memcpy(gt.data() + 3 * T + off, &gr, 2);
    }
  CUDA_TRY(cudaMalloc(&p->gtiles, gt.size()));
  CUDA_TRY(cudaMemcpy(p->gtiles, gt.data(), gt.size(), cudaMemcpyHostToDevice));

  // engine order (32-bit words), per channel h: R rows (c = 0..R-1) of 8192 words; inside a row
  //   w = (cc*128 + k1)*4 + 2*pp + part  holds the bf16 pair (part ? imag : real) of k_f at inner frequencies
  //   k' = k1 + 128*k2 with k2 = 4cc + 2pp and k2 + 1; natural frequency k = c + R*k'.
  // perm[c*8192 + w] = 2 * (natural index of the first element) + part; the second element is 128*R further.
  const int R = p->R;
  std::vector<int> perm(seqlen);
  for (int cr = 0; cr < R; ++cr)
    for (int cc = 0; cc < 16; ++cc)
      for (int k1 = 0; k1 < 128; ++k1)
        for (int pp = 0; pp < 2; ++pp)
          for (int part = 0; part < 2; ++part)
            perm[cr * kInner + (cc * 128 + k1) * 4 + 2 * pp + part] =
                (cr + R * (k1 + 128 * (4 * cc + 2 * pp))) * 2 + part;
  CUDA_TRY(cudaMalloc(&p->perm, perm.size() * sizeof(int)));
  CUDA_TRY(cudaMemcpy(p->perm, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice));

  using namespace bffc::r128;
  CUDA_TRY(cudaFuncSetAttribute(fwd_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  CUDA_TRY(cudaFuncSetAttribute(fwd_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  CUDA_TRY(cudaFuncSetAttribute(fwd_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kSmemTotalGated));
  CUDA_TRY(cudaFuncSetAttribute(fwd_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  CUDA_TRY(cudaFuncSetAttribute(dkf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotalDkf));
  CUDA_TRY(cudaFuncSetAttribute(dkf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotalDkf));
  CUDA_TRY(cudaFuncSetAttribute(dkf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotalDkfGated));
  *out = p;
  return BFFC_OK;
}

int bffc_plan_destroy(bffc_plan* p) {
  if (!p) return BFFC_OK;
  cudaFree(p->dftC);
  cudaFree(p->dftS);
  cudaFree(p->gtiles);
  cudaFree(p->perm);
  delete p;
  return BFFC_OK;
}

int bffc_kf_pack(const bffc_plan* p, const void* kf_natural, void* kf_engine, int H, int conj, void* stream) {
  if (!p || !kf_natural || !kf_engine || H <= 0) return fail(BFFC_ERR_INVALID, "bffc_kf_pack: bad argument");
  dim3 grid(64, H);
  kf_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(kf_natural), static_cast<uint32_t*>(kf_engine), p->perm, p->N, 128 * p->R,
      1.0f / float(p->N), conj);
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

int bffc_dkf_unpack(const bffc_plan* p, const void* dkf_engine, void* dkf_natural, int H, void* stream) {
  if (!p || !dkf_engine || !dkf_natural || H <= 0) return fail(BFFC_ERR_INVALID, "bffc_dkf_unpack: bad argument");
  dim3 grid(64, H);
  bffc::r128::dkf_unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float2*>(dkf_engine), static_cast<float2*>(dkf_natural), p->N, p->R);
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

}  // extern "C"

// Composite sizes keep the outer stage's output as two bf16 planes (real, imaginary) of pairs*H*N elements.
static size_t plane_bytes(const bffc_plan* p, int B, int H) { return size_t((B + 1) / 2) * H * p->N * 2; }

extern "C" size_t bffc_workspace_bytes(const bffc_plan* p, int B, int H, int L) {
  (void)L;
  if (!p || p->R == 1) return 0;
  return 4 * plane_bytes(p, B, H);      // forward uses 2 planes; backward transforms u and dout: 4 planes
}

static int make_map(CUtensorMap* map, const void* base, int rows, int L) {
  // (rows, L) bf16 viewed as [row][L/64][64]; box = one (128 x 64) tile, 128B swizzle;
  // tile rows >= L/64 are out of bounds: zero-filled on load (implicit padding), dropped on store.
  cuuint64_t dims[3] = {64, cuuint64_t(L / 64), cuuint64_t(rows)};
  cuuint64_t strides[2] = {128, cuuint64_t(L) * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
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
  prm.pregate = nullptr;
  prm.postgate = nullptr;
  prm.dbg = nullptr;
  prm.dbg_stages = 0;
}

// fused 8192-point kernel on (B, H, L) real sequences
static int launch_fused(const bffc_plan* p, const void* u, const void* kf, const void* pregate, const void* postgate,
                        void* y, int B, int H, int L, float* dbg, int dbg_stages, int max_units, cudaStream_t st) {
  if (L % 64 != 0) return fail(BFFC_ERR_UNSUPPORTED, "L=%d must be a multiple of 64 for seqlen 8192 in this build", L);
  CUtensorMap tm_u, tm_y, tm_g;
  if (int rc = make_map(&tm_u, u, B * H, L)) return rc;
  if (int rc = make_map(&tm_y, y, B * H, L)) return rc;
  if (int rc = make_map(&tm_g, pregate ? pregate : u, B * H, L)) return rc;
  bffc::FwdParams prm;
  fill_params(p, prm, kf);
  prm.pregate = static_cast<const uint32_t*>(pregate);
  prm.postgate = static_cast<const uint32_t*>(postgate);
  prm.B = B; prm.H = H; prm.L = L;
  prm.pairs = (B + 1) / 2;
  prm.ksteps = (L / 64 + 15) / 16;
  prm.units = H * prm.pairs;
  if (max_units > 0 && prm.units > max_units) prm.units = max_units;
  prm.dbg = dbg;
  prm.dbg_stages = dbg_stages;
  int grid = (prm.units + 1) / 2;
  if (grid > p->num_sms) grid = p->num_sms;
  using namespace bffc::r128;
  if (dbg)
    fwd_kernel<true, false, false><<<grid, kThreads, kSmemTotal, st>>>(tm_u, tm_y, tm_g, prm);
  else if (pregate || postgate)
    fwd_kernel<false, true, false><<<grid, kThreads, kSmemTotalGated, st>>>(tm_u, tm_y, tm_g, prm);
  else
    fwd_kernel<false, false, false><<<grid, kThreads, kSmemTotal, st>>>(tm_u, tm_y, tm_g, prm);
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

// fused kernel on complex rows held in two planes (in place); rows = pairs * kf_rows
static int launch_planes(const bffc_plan* p, void* pre, void* pim, const void* kf, int pairs, int kf_rows, cudaStream_t st) {
  CUtensorMap tm_r, tm_i;
  if (int rc = make_map(&tm_r, pre, pairs * kf_rows, kInner)) return rc;
  if (int rc = make_map(&tm_i, pim, pairs * kf_rows, kInner)) return rc;
  bffc::FwdParams prm;
  fill_params(p, prm, kf);
  prm.B = 2 * pairs; prm.H = kf_rows; prm.L = kInner;
  prm.pairs = pairs;
  prm.ksteps = 8;
  prm.units = kf_rows * pairs;
  int grid = (prm.units + 1) / 2;
  if (grid > p->num_sms) grid = p->num_sms;
  using namespace bffc::r128;
  fwd_kernel<false, false, true><<<grid, kThreads, kSmemTotal, st>>>(tm_r, tm_r, tm_i, prm);
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

template <int R>
static void launch_outer(bool inverse, bool gated, const bffc::outer::OuterParams& op, cudaStream_t st) {
  dim3 grid(bffc::outer::kM / (bffc::outer::kVec * 128), op.H, op.pairs);
  using namespace bffc::outer;
  if (!inverse) {
    if (gated) fwd_kernel<R, true><<<grid, 128, 0, st>>>(op);
    else fwd_kernel<R, false><<<grid, 128, 0, st>>>(op);
  } else {
    if (gated) inv_kernel<R, true><<<grid, 128, 0, st>>>(op);
    else inv_kernel<R, false><<<grid, 128, 0, st>>>(op);
  }
}
static int outer_stage(const bffc_plan* p, bool inverse, bool gated, const bffc::outer::OuterParams& op, cudaStream_t st) {
  switch (p->R) {
    case 2: launch_outer<2>(inverse, gated, op, st); break;
    case 4: launch_outer<4>(inverse, gated, op, st); break;
    case 8: launch_outer<8>(inverse, gated, op, st); break;
    default: return fail(BFFC_ERR_UNSUPPORTED, "outer radix %d not supported", p->R);
  }
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

static int check_common(const bffc_plan* p, int B, int H, int L, const void* a, const void* b, const void* c) {
  if (!p) return fail(BFFC_ERR_INVALID, "null plan");
  if (B <= 0 || H <= 0 || L <= 0 || L > p->N) return fail(BFFC_ERR_INVALID, "bad shape B=%d H=%d L=%d (seqlen %d)", B, H, L, p->N);
  if (L % 8 != 0) return fail(BFFC_ERR_UNSUPPORTED, "L=%d must be a multiple of 8 in this build", L);
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15)
    return fail(BFFC_ERR_INVALID, "device pointers must be 16-byte aligned");
  return 0;
}

// y = postgate * conv(u * pregate, k) for any supported size.  `planes`: workspace for composite sizes.
static int conv_forward(const bffc_plan* p, const void* u, const void* kf, const void* pregate, const void* postgate,
                        void* y, int B, int H, int L, void* planes, cudaStream_t st, int* launches) {
  if (p->R == 1) {
    *launches += 1;
    return launch_fused(p, u, kf, pregate, postgate, y, B, H, L, nullptr, 0, 0, st);
  }
  const int pairs = (B + 1) / 2;
  uint8_t* ws = static_cast<uint8_t*>(planes);
  bffc::outer::OuterParams op;
  op.u = static_cast<const uint4*>(u);
  op.pregate = static_cast<const uint4*>(pregate);
  op.postgate = static_cast<const uint4*>(postgate);
  op.y = static_cast<uint4*>(y);
  op.pre = reinterpret_cast<uint4*>(ws);
  op.pim = reinterpret_cast<uint4*>(ws + plane_bytes(p, B, H));
  op.B = B; op.H = H; op.L = L; op.pairs = pairs;
  if (int rc = outer_stage(p, false, pregate != nullptr, op, st)) return rc;
  if (int rc = launch_planes(p, op.pre, op.pim, kf, pairs, H * p->R, st)) return rc;
  if (int rc = outer_stage(p, true, postgate != nullptr, op, st)) return rc;
  *launches += 3;
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
  if (p->R > 1 && (!workspace || workspace_bytes < 2 * plane_bytes(p, B, H)))
    return fail(BFFC_ERR_INVALID, "bffc_fwd: workspace of %zu bytes required", 2 * plane_bytes(p, B, H));
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
  if (!dout || !u || !kf_conj || !du || !dkf) return fail(BFFC_ERR_INVALID, "bffc_bwd: null pointer");
  if (gated && (!kf || !dpregate || !dpostgate)) return fail(BFFC_ERR_INVALID, "bffc_bwd: gated backward needs kf, dpregate, dpostgate");
  if ((reinterpret_cast<uintptr_t>(pregate) | reinterpret_cast<uintptr_t>(postgate) | reinterpret_cast<uintptr_t>(dpregate) |
       reinterpret_cast<uintptr_t>(dpostgate) | reinterpret_cast<uintptr_t>(kf)) & 15)
    return fail(BFFC_ERR_INVALID, "bffc_bwd: gate pointers must be 16-byte aligned");
  if (int rc = check_common(p, B, H, L, u, du, dout)) return rc;
  if ((reinterpret_cast<uintptr_t>(dkf) | reinterpret_cast<uintptr_t>(kf_conj) | reinterpret_cast<uintptr_t>(workspace)) & 15)
    return fail(BFFC_ERR_INVALID, "bffc_bwd: dkf / kf / workspace must be 16-byte aligned");
  if (p->R > 1 && (!workspace || workspace_bytes < 4 * plane_bytes(p, B, H)))
    return fail(BFFC_ERR_INVALID, "bffc_bwd: workspace of %zu bytes required", 4 * plane_bytes(p, B, H));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int launches = 0;
  // du = corr(dout, k) = circular conv with conj(k_f): the forward path on dout
  // (reference: kernels_bf16/monarch_cuda_32_16_16_bwd_kernel_bf16.h:740-815)
  if (!gated) {
    if (int rc = conv_forward(p, dout, kf_conj, nullptr, nullptr, du, B, H, L, workspace, st, &launches)) return rc;
  } else {
    // y = q * conv(u*p, k)  (conv.py:3856-3939; kernels_bf16/..._bwd_kernel_bf16.h:836-906; host recompute
    // monarch_cuda_interface_bwd_bf16.cu:798-808).  With dx = corr(dout*q, k):
    //   dpostgate = dout * conv(u*p, k),  du = p * dx,  dpregate = u * dx     — three passes of the forward path
    if (int rc = conv_forward(p, u, kf, pregate, dout, dpostgate, B, H, L, workspace, st, &launches)) return rc;
    if (int rc = conv_forward(p, dout, kf_conj, postgate, pregate, du, B, H, L, workspace, st, &launches)) return rc;
    if (int rc = conv_forward(p, dout, kf_conj, postgate, u, dpregate, B, H, L, workspace, st, &launches)) return rc;
  }
  // dk_f = sum_b FFT(dout*q) * conj(FFT(u*p))
  const int pairs = (B + 1) / 2;
  bffc::DkfParams prm;
  prm.dftC = p->dftC;
  prm.dftS = p->dftS;
  prm.gtiles = p->gtiles;
  prm.dkf = static_cast<float2*>(dkf);
  prm.pairs = pairs;
  using namespace bffc::r128;
  if (p->R == 1) {
    if (L % 64 != 0) return fail(BFFC_ERR_UNSUPPORTED, "L=%d must be a multiple of 64 for seqlen 8192 in this build", L);
    CUtensorMap tm_u, tm_d, tm_p, tm_q;
    if (int rc = make_map(&tm_u, u, B * H, L)) return rc;
    if (int rc = make_map(&tm_d, dout, B * H, L)) return rc;
    if (int rc = make_map(&tm_p, gated ? pregate : u, B * H, L)) return rc;
    if (int rc = make_map(&tm_q, gated ? postgate : dout, B * H, L)) return rc;
    prm.B = B; prm.H = H; prm.L = L;
    prm.ksteps = (L / 64 + 15) / 16;
    prm.gated = gated ? 1 : 0;
    int grid = H < p->num_sms ? H : p->num_sms;
    dkf_kernel<false><<<grid, kThreads, gated ? kSmemTotalDkfGated : kSmemTotalDkf, st>>>(tm_u, tm_d, tm_p, tm_q, prm);
    CUDA_TRY(cudaGetLastError());
    launches += 1;
  } else {
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    const size_t pb = plane_bytes(p, B, H);
    bffc::outer::OuterParams ou, od;
    ou.u = static_cast<const uint4*>(u); ou.pregate = static_cast<const uint4*>(pregate); ou.postgate = nullptr; ou.y = nullptr;
    ou.pre = reinterpret_cast<uint4*>(ws); ou.pim = reinterpret_cast<uint4*>(ws + pb);
    ou.B = B; ou.H = H; ou.L = L; ou.pairs = pairs;
    od = ou;
    od.u = static_cast<const uint4*>(dout);
    od.pregate = static_cast<const uint4*>(postgate);
    od.pre = reinterpret_cast<uint4*>(ws + 2 * pb); od.pim = reinterpret_cast<uint4*>(ws + 3 * pb);
    if (int rc = outer_stage(p, false, gated, ou, st)) return rc;
    if (int rc = outer_stage(p, false, gated, od, st)) return rc;
    const int rows = H * p->R;
    CUtensorMap tur, tui, tdr, tdi;
    if (int rc = make_map(&tur, ou.pre, pairs * rows, kInner)) return rc;
    if (int rc = make_map(&tui, ou.pim, pairs * rows, kInner)) return rc;
    if (int rc = make_map(&tdr, od.pre, pairs * rows, kInner)) return rc;
    if (int rc = make_map(&tdi, od.pim, pairs * rows, kInner)) return rc;
    prm.B = 2 * pairs; prm.H = rows; prm.L = kInner;
    prm.ksteps = 8;
    prm.gated = 0;
    int grid = rows < p->num_sms ? rows : p->num_sms;
    dkf_kernel<true><<<grid, kThreads, kSmemTotalDkf, st>>>(tur, tdr, tui, tdi, prm);
    CUDA_TRY(cudaGetLastError());
    launches += 3;
  }
  g_launches = launches;
  return BFFC_OK;
}

int bffc_debug_fwd_stages(const bffc_plan* p, const void* u, const void* kf, void* y, int B, int H, int L, float* dump,
                          int max_stages, void* stream) {
  if (!dump || max_stages <= 0 || !p || p->R != 1) return -BFFC_ERR_INVALID;
  if (check_common(p, B, H, L, u, y, kf)) return -BFFC_ERR_INVALID;
  int rc = launch_fused(p, u, kf, nullptr, nullptr, y, B, H, L, dump, max_stages, 1, static_cast<cudaStream_t>(stream));
  if (rc) return -rc;
  return max_stages < 4 ? max_stages : 4;
}

}  // extern "C"

// bffc.cu — host side of the C ABI declared in include/bffc.h (plan tables, TMA descriptors, launches).
//
// Replaces, for the fused FFT-convolution path only:
//   FlashFFTConv.__init__ tables            (reference flashfftconv/conv.py:72-551)
//   k_f permutation + cast per call         (conv.py:640, :676, :1423-1424)
//   pybind entry + C++ dispatch + launcher  (csrc/flashfftconv/monarch.cpp:16-56,
//                                            monarch_cuda/monarch_fwd.h:296-376,
//                                            monarch_cuda_interface_fwd_bf16.cu:656-760)
// No torch types cross this boundary; there is no CPU fallback.
#include "bffc.h"
#include "fwd_r128.cuh"
#include "dkf_r128.cuh"

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

}  // namespace

struct bffc_plan {
  int N;
  int dtype;
  int device;
  __nv_bfloat16* dftC = nullptr;
  __nv_bfloat16* dftS = nullptr;
  uint8_t* gtiles = nullptr;
  int* perm = nullptr;  // engine index -> natural frequency index
  int num_sms = 0;
};

extern "C" {

int bffc_abi_version(void) { return BFFC_ABI_VERSION; }
const char* bffc_last_error(void) { return g_err; }
int bffc_last_launch_count(void) { return g_launches; }

int bffc_supported(int seqlen, int dtype) { return (seqlen == 8192 && dtype == BFFC_DTYPE_BF16) ? 1 : 0; }

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
  p->dtype = dtype;
  CUDA_TRY(cudaGetDevice(&p->device));
  CUDA_TRY(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));

  const double PI = 3.14159265358979323846;
  // outer radix-128 DFT, cos / sin planes (symmetric, K-major rows)
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
  // (the 128B swizzle TMA / UMMA use), planes Gr then Gi.
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

  // engine order (32-bit words): w = (c*128 + k1)*4 + 2*pp + part holds the bf16 pair
  //   (part ? imag : real) of k_f at k2 = 4c + 2pp and k2 + 1, natural frequency k = k1 + 128*k2.
  // perm[w] = natural index of the first element * 2 + part; the second element is 128 further.
  std::vector<int> perm(seqlen);
  for (int c = 0; c < 16; ++c)
    for (int k1 = 0; k1 < 128; ++k1)
      for (int pp = 0; pp < 2; ++pp)
        for (int part = 0; part < 2; ++part)
          perm[(c * 128 + k1) * 4 + 2 * pp + part] = (k1 + 128 * (4 * c + 2 * pp)) * 2 + part;
  CUDA_TRY(cudaMalloc(&p->perm, perm.size() * sizeof(int)));
  CUDA_TRY(cudaMemcpy(p->perm, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice));

  CUDA_TRY(cudaFuncSetAttribute(bffc::r128::fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                bffc::r128::kSmemTotal));
  CUDA_TRY(cudaFuncSetAttribute(bffc::r128::fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                bffc::r128::kSmemTotal));
  CUDA_TRY(cudaFuncSetAttribute(bffc::r128::fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                bffc::r128::kSmemTotalGated));
  CUDA_TRY(cudaFuncSetAttribute(bffc::r128::dkf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                bffc::r128::kSmemTotalDkf));
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
  dim3 grid((p->N + 255) / 256 > 64 ? 64 : (p->N + 255) / 256, H);
  kf_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(kf_natural), static_cast<uint32_t*>(kf_engine), p->perm, p->N, 128, 1.0f / float(p->N),
      conj);
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

int bffc_dkf_unpack(const bffc_plan* p, const void* dkf_engine, void* dkf_natural, int H, void* stream) {
  if (!p || !dkf_engine || !dkf_natural || H <= 0) return fail(BFFC_ERR_INVALID, "bffc_dkf_unpack: bad argument");
  dim3 grid(32, H);
  bffc::r128::dkf_unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float2*>(dkf_engine), static_cast<float2*>(dkf_natural), p->N);
  CUDA_TRY(cudaGetLastError());
  return BFFC_OK;
}

size_t bffc_workspace_bytes(const bffc_plan*, int, int, int) { return 0; }

static int make_map(CUtensorMap* map, const void* base, int BH, int L) {
  // (B*H, L) bf16 viewed as [seq][row = L/64][col = 64]; box = one (128 x 64) tile, 128B swizzle;
  // rows >= L/64 are out of bounds: zero-filled on load (implicit padding), dropped on store.
  cuuint64_t dims[3] = {64, cuuint64_t(L / 64), cuuint64_t(BH)};
  cuuint64_t strides[2] = {128, cuuint64_t(L) * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BFFC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(r));
  return 0;
}

static int launch_fwd(const bffc_plan* p, const void* u, const void* kf, const void* pregate, const void* postgate,
                      void* y, int B, int H, int L, float* dbg, int dbg_stages, int max_units, void* stream) {
  if (!p || !u || !kf || !y) return fail(BFFC_ERR_INVALID, "bffc_fwd: null pointer");
  if (B <= 0 || H <= 0 || L <= 0 || L > p->N) return fail(BFFC_ERR_INVALID, "bffc_fwd: bad shape B=%d H=%d L=%d", B, H, L);
  if (L % 64 != 0) return fail(BFFC_ERR_UNSUPPORTED, "bffc_fwd: L=%d must be a multiple of 64 in this build", L);
  if ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(kf)) & 15)
    return fail(BFFC_ERR_INVALID, "bffc_fwd: u, y and kf must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(pregate) | reinterpret_cast<uintptr_t>(postgate)) & 15)
    return fail(BFFC_ERR_INVALID, "bffc_fwd: gates must be 16-byte aligned");
  CUtensorMap tm_u, tm_y, tm_g;
  if (int rc = make_map(&tm_u, u, B * H, L)) return rc;
  if (int rc = make_map(&tm_y, y, B * H, L)) return rc;
  if (int rc = make_map(&tm_g, pregate ? pregate : u, B * H, L)) return rc;
  bffc::FwdParams prm;
  prm.kf = static_cast<const uint32_t*>(kf);
  prm.dftC = p->dftC;
  prm.dftS = p->dftS;
  prm.gtiles = p->gtiles;
  prm.pregate = static_cast<const uint32_t*>(pregate);
  prm.postgate = static_cast<const uint32_t*>(postgate);
  prm.L = L;
  prm.B = B;
  prm.H = H;
  prm.pairs = (B + 1) / 2;
  prm.ksteps = (L / 64 + 15) / 16;
  prm.units = H * prm.pairs;
  if (max_units > 0 && prm.units > max_units) prm.units = max_units;
  prm.dbg = dbg;
  prm.dbg_stages = dbg_stages;
  int grid = (prm.units + 1) / 2;
  if (grid > p->num_sms) grid = p->num_sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  using namespace bffc::r128;
  if (dbg)
    fwd_kernel<true, false><<<grid, kThreads, kSmemTotal, st>>>(tm_u, tm_y, tm_g, prm);
  else if (pregate)
    fwd_kernel<false, true><<<grid, kThreads, kSmemTotalGated, st>>>(tm_u, tm_y, tm_g, prm);
  else
    fwd_kernel<false, false><<<grid, kThreads, kSmemTotal, st>>>(tm_u, tm_y, tm_g, prm);
  CUDA_TRY(cudaGetLastError());
  g_launches = 1;
  return BFFC_OK;
}

int bffc_fwd(const bffc_plan* p, const void* u, const void* kf, const void* pregate, const void* postgate, void* y,
             int B, int H, int L, void*, size_t, void* stream) {
  if ((pregate == nullptr) != (postgate == nullptr))
    return fail(BFFC_ERR_INVALID, "bffc_fwd: pregate and postgate must both be given or both be null");
  return launch_fwd(p, u, kf, pregate, postgate, y, B, H, L, nullptr, 0, 0, stream);
}

int bffc_bwd(const bffc_plan* p, const void* dout, const void* u, const void* kf, const void* kf_conj,
             const void* pregate, const void* postgate, void* du, void* dkf, void* dpregate, void* dpostgate, int B, int H,
             int L, void*, size_t, void* stream) {
  if ((pregate == nullptr) != (postgate == nullptr))
    return fail(BFFC_ERR_INVALID, "bffc_bwd: pregate and postgate must both be given or both be null");
  if (pregate) return fail(BFFC_ERR_UNSUPPORTED, "bffc_bwd: gated backward not implemented yet");
  (void)kf; (void)dpregate; (void)dpostgate;
  if (!p || !dout || !u || !kf_conj || !du || !dkf) return fail(BFFC_ERR_INVALID, "bffc_bwd: null pointer");
  // du = corr(dout, k) = circular conv with conj(k_f): the forward kernel on dout (kernels_bf16/..._bwd_kernel_bf16.h:740-815)
  if (int rc = launch_fwd(p, dout, kf_conj, nullptr, nullptr, du, B, H, L, nullptr, 0, 0, stream)) return rc;
  // dk_f = sum_b FFT(dout) * conj(FFT(u))
  if ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(dkf)) & 15)
    return fail(BFFC_ERR_INVALID, "bffc_bwd: u and dkf must be 16-byte aligned");
  CUtensorMap tm_u, tm_d;
  if (int rc = make_map(&tm_u, u, B * H, L)) return rc;
  if (int rc = make_map(&tm_d, dout, B * H, L)) return rc;
  bffc::DkfParams prm;
  prm.dftC = p->dftC;
  prm.dftS = p->dftS;
  prm.gtiles = p->gtiles;
  prm.dkf = static_cast<float2*>(dkf);
  prm.B = B; prm.H = H; prm.L = L;
  prm.pairs = (B + 1) / 2;
  prm.ksteps = (L / 64 + 15) / 16;
  int grid = H < p->num_sms ? H : p->num_sms;
  bffc::r128::dkf_kernel<<<grid, bffc::r128::kThreads, bffc::r128::kSmemTotalDkf, static_cast<cudaStream_t>(stream)>>>(
      tm_u, tm_d, prm);
  CUDA_TRY(cudaGetLastError());
  g_launches = 2;
  return BFFC_OK;
}

int bffc_debug_fwd_stages(const bffc_plan* p, const void* u, const void* kf, void* y, int B, int H, int L, float* dump,
                          int max_stages, void* stream) {
  if (!dump || max_stages <= 0) return -BFFC_ERR_INVALID;
  int rc = launch_fwd(p, u, kf, nullptr, nullptr, y, B, H, L, dump, max_stages, 1, stream);
  if (rc) return -rc;
  return max_stages < 4 ? max_stages : 4;
}

}  // extern "C"

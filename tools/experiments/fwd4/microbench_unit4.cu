// Micro-benchmark for the flop-lean inner kernel (fwd4): the tcgen05 MMA sequence of one unit when every stage takes the
// data as the A operand from shared memory (SS) and a small DFT matrix as B:
//   stage radix 16: 4 chunks x (re, im) = 8 MMAs  M=128 N=32 K=16, A MN-major 128B-swizzled (4 KB per MMA)
//   stage radix 32: 2 chunks x (re, im) x 2 K steps = 8 MMAs  N=64
//   inverse radix 32 (row local): K-major A, 8 MMAs N=64
// per unit: 16,16,32 forward and 32,16,16 inverse = 48 MMAs, tensor model 8*(16+16+32)*2 = 1024 cycles if the pipe is
// compute bound, 6 x 256 = 1536 if every MMA is bound by its 4 KB A read at 128 B/clk.  1..4 issuing warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flash-fft-conv_b200/csrc -o gpurun_out/mbu4 tools/microbench_unit4.cu
#include "ptx.cuh"
#include <cstdio>
using namespace bffc;

constexpr int kSlot = 32768;
// instruction descriptor with both major bits selectable
__host__ __device__ constexpr uint32_t idesc4(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) | (uint32_t(n >> 3) << 17) |
         (uint32_t(128 >> 4) << 24);
}
__device__ __forceinline__ uint64_t a_mn_desc(uint32_t s, uint32_t lbo) { return make_sdesc(s, lbo, 1024, 2); }   // MN-major, SW128
__device__ __forceinline__ uint64_t a_k_desc(uint32_t s) { return make_sdesc(s, 16, 1024, 2); }                   // K-major, SW128
__device__ __forceinline__ uint64_t b_k_desc(uint32_t s) { return make_sdesc(s, 128, 256, 0); }                   // K-major, no swizzle

__global__ void __launch_bounds__(512, 1) k_unit4(int npipes, int units, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bars[4];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, pipe = tid >> 7;
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
  if (tid < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  for (int i = tid; i < (4 * kSlot + 16384) / 4; i += 512) reinterpret_cast<uint32_t*>(smem_raw + (sb - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t ID32 = idesc4(32, true, false), ID64 = idesc4(64, true, false), ID64K = idesc4(64, false, false);
  const uint32_t tD0 = tptr + 128 * pipe;
  const uint32_t sX = sb + pipe * kSlot, sB = sb + 4 * kSlot, bar = smem_u32(&bars[pipe]);
  long long t0 = clock64();
  if ((tid & 127) < 32 && pipe < npipes) {
    if (elect_one()) {
      uint32_t phase = 0;
      for (int u = 0; u < units; ++u) {
        for (int dir = 0; dir < 2; ++dir) {
          if (mode & 1) {                                   // two radix-16 stages (MN-major A)
            for (int st = 0; st < 2; ++st) {
              for (int c = 0; c < 4; ++c)
                for (int pl = 0; pl < 2; ++pl)
                  mma_ss(tD0 + 32 * c, a_mn_desc(sX + pl * 16384 + c * 4096, 2048), b_k_desc(sB + (2 * c + pl) * 1024), ID32, pl);
              mma_commit(bar); mbar_wait(bar, phase); phase ^= 1;
            }
          }
          if (mode & 2) {                                   // radix-32 stage
            for (int c = 0; c < 2; ++c)
              for (int pl = 0; pl < 2; ++pl)
                for (int ks = 0; ks < 2; ++ks) {
                  if (dir == 0) mma_ss(tD0 + 64 * c, a_mn_desc(sX + pl * 16384 + c * 8192 + ks * 2048, 4096), b_k_desc(sB + 8192 + (2 * pl + ks) * 2048), ID64, pl | ks);
                  else mma_ss(tD0 + 64 * c, a_k_desc(sX + pl * 16384 + 64 * c + 32 * ks), b_k_desc(sB + 8192 + (2 * pl + ks) * 2048), ID64K, pl | ks);
                }
            mma_commit(bar); mbar_wait(bar, phase); phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (tid < 32) tmem_dealloc(tptr, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h;
  const int smem = 4 * kSlot + 16384 + 2048;
  cudaFuncSetAttribute(k_unit4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int units = 50;
  for (int mode = 1; mode <= 3; ++mode)
    for (int np = 1; np <= 4; ++np) {
      k_unit4<<<1, 512, smem>>>(np, units, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const int model = ((mode & 1) ? 4 * 128 : 0) + ((mode & 2) ? 2 * 256 : 0);
      printf("mode=%d (%s) pipelines=%d  cycles per unit and pipeline %7.0f  per unit on the SM %7.0f  (flop model %d, A-read model %d)  %s\n", mode,
             mode == 1 ? "4 radix-16 stages" : mode == 2 ? "2 radix-32 stages" : "all six stages", np, double(h) / units, double(h) / units / np, model,
             ((mode & 1) ? 4 * 256 : 0) + ((mode & 2) ? 2 * 256 : 0), cudaGetErrorString(e));
    }
  return 0;
}

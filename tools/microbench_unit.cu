// Micro-benchmark: the exact tcgen05 MMA sequence of one fwd3 unit (stage 1: 24 TS MMAs, stage 2/3: 8 SS MMAs each,
// stage 4: 24 TS MMAs), commit + mbarrier wait after every stage, issued by 1..3 independent warps ("pipelines") that
// share the SM's tensor pipe — no loads, no passes.  Answers: how many cycles per unit does the tensor pipe need?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flash-fft-conv_b200/csrc -o gpurun_out/mbu tools/microbench_unit.cu
#include "ptx.cuh"
#include <cstdio>
using namespace bffc;

constexpr int kTile = 16384, kSlot = 32768;
__device__ __forceinline__ uint64_t tile_desc(uint32_t s) { return make_sdesc(s, kTile, 1024, 2); }
__device__ __forceinline__ uint64_t pair_desc(uint32_t s, uint32_t lbo) { return make_sdesc(s, lbo, 1024, 2); }
__device__ __forceinline__ uint64_t atile_desc(uint32_t s) { return make_sdesc(s, 16, 1024, 2); }

// mode bit0: stage 1/4 (TS) on, bit1: stage 2/3 (SS) on; waits: 1 = wait after every stage, 0 = only at the end of a unit
__global__ void __launch_bounds__(384, 1) k_unit(int npipes, int units, int mode, int waits, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bars[3];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, pipe = tid >> 7;
  if (tid == 0) { for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
  if (tid < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  for (int i = tid; i < (3 * kSlot + 32768) / 4; i += 384) reinterpret_cast<uint32_t*>(smem_raw + (sb - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t ID128 = make_idesc(1, 128, true, false), ID64 = make_idesc(1, 64, true, false);
  const uint32_t tC0 = tptr, tS0 = tptr + 64, tD0 = tptr + 128 + 128 * pipe;
  const uint32_t sX = sb + pipe * kSlot, sG0 = sb + 3 * kSlot, bar = smem_u32(&bars[pipe]);
  long long t0 = clock64();
  if ((tid & 127) < 32 && pipe < npipes) {
    if (elect_one()) {
      uint32_t phase = 0;
      for (int u = 0; u < units; ++u) {
        if (mode & 1) {
          for (int s = 0; s < 8; ++s) mma_ts(tD0, tC0 + 8 * s, tile_desc(sX + s * 2048), ID128, s > 0);
          for (int s = 0; s < 8; ++s) mma_ts(tD0, tS0 + 8 * s, tile_desc(sX + kTile + s * 2048), ID64, 1);
          for (int s = 0; s < 8; ++s) mma_ts(tD0 + 64, tS0 + 8 * s, tile_desc(sX + s * 2048), ID64, 1);
          if (waits) { mma_commit(bar); mbar_wait(bar, phase); phase ^= 1; }
        }
        if (mode & 2) {
          for (int rep = 0; rep < 2; ++rep) {
            for (int s = 0; s < 4; ++s) mma_ss(tD0, atile_desc(sX + 32 * s), pair_desc(sG0 + s * 2048, 8192), ID128, s > 0);
            for (int s = 0; s < 4; ++s) mma_ss(tD0, atile_desc(sX + kTile + 32 * s), pair_desc(sG0 + 16384 + s * 2048, 8192), ID128, 1);
            if (waits) { mma_commit(bar); mbar_wait(bar, phase); phase ^= 1; }
          }
        }
        if (mode & 1) {
          for (int s = 0; s < 8; ++s) mma_ts(tD0, tC0 + 8 * s, tile_desc(sX + s * 2048), ID128, s > 0);
          for (int s = 0; s < 8; ++s) mma_ts(tD0, tS0 + 8 * s, tile_desc(sX + kTile + s * 2048), ID64, 1);
          for (int s = 0; s < 8; ++s) mma_ts(tD0 + 64, tS0 + 8 * s, tile_desc(sX + s * 2048), ID64, 1);
          if (waits) { mma_commit(bar); mbar_wait(bar, phase); phase ^= 1; }
        }
        if (!waits) { mma_commit(bar); mbar_wait(bar, phase); phase ^= 1; }
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
  cudaFuncSetAttribute(k_unit, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kSlot + 32768 + 2048);
  const int units = 50;
  for (int waits = 1; waits >= 0; --waits)
    for (int mode = 1; mode <= 3; ++mode)
      for (int np = 1; np <= 3; ++np) {
        k_unit<<<1, 384, 3 * kSlot + 32768 + 2048>>>(np, units, mode, waits, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const int model = ((mode & 1) ? 2048 : 0) + ((mode & 2) ? 1024 : 0);
        printf("waits=%d mode=%d (%s) pipelines=%d  cycles per unit and pipeline %7.0f  per unit on the SM %7.0f  (tensor model %d)  %s\n", waits, mode,
               mode == 1 ? "TS stages 1+4" : mode == 2 ? "SS stages 2+3" : "all four stages", np, double(h) / units, double(h) / units / np, model,
               cudaGetErrorString(e));
      }
  return 0;
}

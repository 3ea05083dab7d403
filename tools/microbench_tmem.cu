// Micro-benchmarks that pin the constants the kernel design depends on (B200, sm_100a):
//   1. tcgen05.ld (TMEM -> registers) throughput per SM vs number of warps and vector width
//   2. tcgen05.st throughput
//   3. latency of one tcgen05.ld + wait
//   4. MMA batch latency: issue K MMAs (N=128, A from TMEM) + commit + mbarrier wait, from one thread
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flash-fft-conv_b200/csrc -o gpurun_out/mb tools/microbench_tmem.cu
#include "ptx.cuh"
#include <cstdio>
using namespace bffc;

__global__ void __launch_bounds__(512, 1) k_ld(int iters, int mode, long long* out) {
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr + (uint32_t((threadIdx.x >> 5) & 3) * 32 << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {            // x16 loads
    for (int i = 0; i < iters; ++i) {
      uint32_t v[16];
      tmem_ld16(tb + ((i * 16) & 255), v);
      tmem_ld_wait();
      reg_fence(v);
      acc += v[0] + v[15];
    }
  } else if (mode == 1) {     // x32 loads
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld32(tb + ((i * 32) & 255), v);
      tmem_ld_wait();
      reg_fence(v);
      acc += v[0] + v[31];
    }
  } else if (mode == 2) {     // 4 x16 loads in flight, one wait
    for (int i = 0; i < iters; ++i) {
      uint32_t a[16], b[16], c[16], d[16];
      tmem_ld16(tb + 0, a); tmem_ld16(tb + 16, b); tmem_ld16(tb + 32, c); tmem_ld16(tb + 48, d);
      tmem_ld_wait();
      reg_fence(a); reg_fence(b); reg_fence(c); reg_fence(d);
      acc += a[0] + b[1] + c[2] + d[3];
    }
  } else if (mode == 3) {     // x16 stores
    for (int i = 0; i < iters; ++i) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = acc + j + i;
      tmem_st16(tb + ((i * 16) & 255), v);
      tmem_st_wait();
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678) out[1] = acc;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tptr, 512);
  (void)bar;
}

// one thread issues `nmma` MMAs (M=128, N=ncols, K=16, A from TMEM, B from smem) then commit; measures cycles until the
// mbarrier flips.  B contents irrelevant.
__global__ void __launch_bounds__(128, 1) k_mma(int nmma, int ncols, int reps, int variant, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (sb - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t idesc = make_idesc(1, ncols, true, false);
  long long total = 0;
  uint32_t phase = 0;
  if (threadIdx.x < 32) {
   if (elect_one()) {      // same warp-uniform election pattern as the kernels (operands stay in uniform registers)
    for (int r = 0; r < reps; ++r) {
      long long t0 = clock64();
      if (variant == 0) {          // one dependent accumulation chain
        for (int s = 0; s < nmma; ++s)
          mma_ts(tptr + 128, tptr + 8 * (s & 7), make_sdesc(sb + (s & 7) * 2048, 16384, 1024, 2), idesc, s > 0);
      } else if (variant == 1) {   // two independent chains, interleaved
        for (int s = 0; s < nmma; ++s)
          mma_ts(tptr + 128 + 128 * (s & 1), tptr + 8 * (s & 7), make_sdesc(sb + (s & 7) * 2048, 16384, 1024, 2), idesc, s > 1);
      } else if (variant == 2) {   // three independent chains
        for (int s = 0; s < nmma; ++s)
          mma_ts(tptr + 128 + 128 * (s % 3), tptr + 8 * (s & 7), make_sdesc(sb + (s & 7) * 2048, 16384, 1024, 2), idesc, s > 2);
      } else {                     // one chain, descriptor arithmetic hoisted (constant descriptor / A address)
        const uint64_t d0 = make_sdesc(sb, 16384, 1024, 2);
        const uint32_t a0 = tptr, dd = tptr + 128;
#pragma unroll 8
        for (int s = 0; s < nmma; ++s) mma_ts(dd, a0, d0, idesc, 1);
      }
      mma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      total += clock64() - t0;
    }
    out[0] = total / reps;
   }
   __syncwarp();
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tptr, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h[2];
  const int iters = 2000;
  const char* names[4] = {"ld x16 + wait", "ld x32 + wait", "4 x ld x16, one wait", "st x16 + wait"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      k_ld<<<1, warps * 32, 0>>>(iters, mode, d);
      cudaDeviceSynchronize();
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const double cols = (mode == 1 ? 32.0 : mode == 2 ? 64.0 : 16.0);
      const double bytes = double(iters) * cols * 4 * 32 * warps;
      printf("%-22s warps=%2d  cycles/iter=%7.1f  bytes/clk/SM=%7.1f\n", names[mode], warps, double(h[0]) / iters, bytes / double(h[0]));
    }
  cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  for (int variant = 0; variant < 4; ++variant)
    for (int ncols : {64, 128, 256})
      for (int nmma : {1, 8, 24}) {
        if (ncols == 256 && variant != 0 && variant != 3) continue;   // two 256-column tiles do not fit next to A
        k_mma<<<1, 128, 100000>>>(nmma, ncols, 50, variant, d);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
        printf("mma variant=%d N=%3d count=%2d  cycles issue->barrier=%6lld  (model %d)\n", variant, ncols, nmma, h[0], nmma * ncols / 2);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

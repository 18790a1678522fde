// Microbenchmark 2: what slows tcgen05.mma (M=128, K=16) inside a real kernel?
//   mode bit 0: operands walk over 7 A slots (17408 B, shifts 0..2) and 3 B tiles (24 KB) instead of one fixed tile
//   mode bit 1: 4 other warps run tcgen05.ld (32x32b.x32) loops on other TMEM columns
//   mode bit 2: 4 other warps stream st.shared.v4 into a 32 KB region
//   mode bit 3: a tcgen05.commit + mbarrier wait round trip every 12 MMAs (what the conv kernel does per row)
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../deepdenoiser_b200/csrc/dd_ptx.cuh"
using namespace dd;

__global__ void __launch_bounds__(288, 1) bench(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < 220 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); stop = 0; }
  if (threadIdx.x < 32) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      const uint64_t tmpl = make_desc_sw128(0, 0);
      const uint32_t idesc = make_idesc_f16(128, N);
      uint32_t phase = 0;
      long long t0 = clock64();
      for (int i = 0; i < iters; i += 12) {
        const int row = (mode & 1) ? (i / 12) % 7 : 0;
        for (int s = 0; s < 3; ++s) {
          const uint32_t a0 = base + row * 17408 + ((mode & 1) ? s * 128 : 0);
          const uint32_t b0 = base + 128 * 1024 + ((mode & 1) ? s * 24576 : 0);
          const uint64_t ad = tmpl + (a0 >> 4), bd = tmpl + (b0 >> 4);
          const uint32_t d = tmem + (((i / 12) & 1) ? N : 0);
          umma_f16(d, ad, bd, idesc, 1u);
          umma_f16(d, ad + 2, bd + 2, idesc, 1u);
          umma_f16(d, ad + 4, bd + 4, idesc, 1u);
          umma_f16(d, ad + 6, bd + 6, idesc, 1u);
        }
        if (mode & 8) {
          umma_commit(&bar[0]);
          mbar_wait(&bar[0], phase); phase ^= 1;
          tc_fence_after();
        }
      }
      long long t1 = clock64();
      umma_commit(&bar[1]);
      mbar_wait(&bar[1], 0);
      long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
      stop = 1;
    }
  } else if (warp >= 1 && warp <= 4) {
    if (mode & 2) {
      uint32_t v[32]; uint32_t acc = 0;
      const uint32_t taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 384;
      while (!stop) {
        tmem_ld_32x32(taddr, v); tmem_ld_wait();
        acc += v[0] + v[31];
        tmem_ld_32x32(taddr + 32, v); tmem_ld_wait();
        acc += v[5];
      }
      if (acc == 0x12345678) out[3] = acc;
    }
  } else if (warp >= 5) {
    if (mode & 4) {
      uint4* dst = reinterpret_cast<uint4*>(smem + 180 * 1024);
      int k = 0;
      while (!stop) {
        dst[(k * 128 + (threadIdx.x - 160)) & 2047] = make_uint4(k, k, k, k);
        ++k;
      }
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
  const int iters = 1200;
  for (int N : {192, 64}) {
    for (int mode = 0; mode < 16; ++mode) {
      bench<<<148, 288, 222 * 1024>>>(N, mode, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("N %3d mode %2d [%s%s%s%s]: issue %.1f complete %.1f cyc/mma %s\n", N, mode, (mode & 1) ? "walk " : "", (mode & 2) ? "ldtm " : "",
             (mode & 4) ? "sts " : "", (mode & 8) ? "commit+wait/12" : "", double(h[0]) / iters, double(h[1]) / iters,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}

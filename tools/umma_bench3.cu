// Microbenchmark 3: issue-side cost of tcgen05.mma / tcgen05.commit for the single issuing thread.
//   per iteration: K UMMAs (M=128, K=16, given N) then C commits onto distinct mbarriers (nobody waits on them).
//   Reports cycles per iteration seen by the issuing thread (issue) and until everything completed (complete).
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../deepdenoiser_b200/csrc/dd_ptx.cuh"
using namespace dd;

__global__ void __launch_bounds__(128, 1) bench(int N, int K, int C, int iters, int waitmode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar[8];
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint64_t tmpl = make_desc_sw128(0, 0);
    const uint32_t idesc = make_idesc_f16(128, N);
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t a0 = base + (i % 7) * 17408;
      const uint32_t b0 = base + 128 * 1024;
      const uint64_t ad = tmpl + (a0 >> 4), bd = tmpl + (b0 >> 4);
      const uint32_t d = tmem + ((i & 1) ? 256 : 0);
      for (int k = 0; k < K; ++k) umma_f16(d, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1u);
      for (int c = 0; c < C; ++c) umma_commit(&bar[c]);
      if (waitmode == 1 && C > 0 && i >= 4) {   // lagging wait: the commit issued 4 iterations ago (never blocks long)
        // parity of bar[0] after (i-3) completed phases
        mbar_wait(&bar[0], static_cast<uint32_t>(i - 4) & 1u);
      }
      (void)phase;
    }
    long long t1 = clock64();
    umma_commit(&bar[7]);
    mbar_wait(&bar[7], 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 400;
  for (int N : {32, 96, 192}) {
    for (int K : {2, 4, 7, 12}) {
      for (int C : {0, 1, 2, 3}) {
        bench<<<148, 128, 200 * 1024>>>(N, K, C, iters, 0, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("N %3d K %2d commits %d: issue %.1f complete %.1f cyc/iter %s\n", N, K, C, double(h[0]) / iters, double(h[1]) / iters,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  }
  return 0;
}

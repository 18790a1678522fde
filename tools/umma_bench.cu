// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SW128 K-major operands from smem)
// for different N, dependent vs. alternating accumulators, aligned vs. shifted A.   nvcc -arch=sm_100a
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../deepdenoiser_b200/csrc/dd_ptx.cuh"
using namespace dd;

__global__ void __launch_bounds__(128, 1) bench(int N, int n_acc, int a_shift, int iters, long long* out, int b_rows_stride_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint64_t tmpl = make_desc_sw128(0, 0);
    const uint32_t a0 = base + a_shift * 128, b0 = base + 32 * 1024;
    const uint32_t idesc = make_idesc_f16(128, N);
    long long t0 = clock64();
    const uint64_t ad = tmpl + (a0 >> 4), bd = tmpl + (b0 >> 4);
    const uint32_t d1 = tmem + (n_acc > 1 ? N : 0);
    for (int i = 0; i < iters; i += 8) {
      umma_f16(tmem, ad, bd, idesc, 1u);
      umma_f16(d1, ad + 2, bd + 2, idesc, 1u);
      umma_f16(tmem, ad + 4, bd + 4, idesc, 1u);
      umma_f16(d1, ad + 6, bd + 6, idesc, 1u);
      umma_f16(tmem, ad, bd, idesc, 1u);
      umma_f16(d1, ad + 2, bd + 2, idesc, 1u);
      umma_f16(tmem, ad + 4, bd + 4, idesc, 1u);
      umma_f16(d1, ad + 6, bd + 6, idesc, 1u);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2048;
  for (int grid : {1, 148}) {
    for (int N : {32, 64, 96, 128, 192, 256}) {
      for (int n_acc : {1, 2}) {
        if (n_acc * N > 512) continue;
        for (int shift : {0, 1}) {
          bench<<<grid, 128, 100 * 1024>>>(N, n_acc, shift, iters, d, 0);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("grid %3d N %3d accs %d a_shift %d : issue %.1f cyc/mma, complete %.1f cyc/mma (%.0f MAC/clk) %s\n", grid, N, n_acc,
                 shift, double(h[0]) / iters, double(h[1]) / iters, 128.0 * N * 16 * iters / double(h[1]), e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
    }
  }
  return 0;
}

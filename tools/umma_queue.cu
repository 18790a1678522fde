// Microbenchmark: how many tcgen05.mma can the issuing thread have outstanding?  A burst of n UMMAs (M=128, N=192, K=16) is
// issued on an idle tensor pipe; the thread-side time of the burst (issue) is compared with its completion time.  A deep queue
// shows issue << complete for small n; a queue of depth D throttles the issue to the pipe rate after D instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_queue.bin tools/umma_queue.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../deepdenoiser_b200/csrc/dd_ptx.cuh"
using namespace dd;

template <int NB>
__global__ void __launch_bounds__(128, 1) bench(int N, long long* out) {
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
    const uint32_t a0 = base, b0 = base + 32 * 1024;
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t ad = tmpl + (a0 >> 4), bd = tmpl + (b0 >> 4);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {          // the last repetition is reported (warm instruction cache)
      long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < NB; ++i) umma_f16(tmem + (i & 1) * N, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, 1u);
      long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int NB>
void run(int N, long long* d) {
  cudaFuncSetAttribute(bench<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  bench<NB><<<1, 128, 100 * 1024>>>(N, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N %3d burst %2d : issue %5lld cycles (%.1f / mma), complete %5lld cycles (%.1f / mma) %s\n", N, NB, h[0], double(h[0]) / NB, h[1],
         double(h[1]) / NB, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  for (int N : {64, 192, 256}) {
    run<1>(N, d); run<2>(N, d); run<3>(N, d); run<4>(N, d); run<6>(N, d); run<8>(N, d); run<12>(N, d); run<16>(N, d); run<24>(N, d); run<32>(N, d);
  }
  return 0;
}

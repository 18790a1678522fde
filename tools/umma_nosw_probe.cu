// Hardware probe for the shared-memory operand layout of the fused compose kernel (csrc/compose_rows.cuh):
// K-major, NO swizzle ("interleaved" core matrices): element (row m, 16-byte k-chunk j) at  base + j*LBO + m*16
// when SBO (the stride between 8-row groups) is 128 B, i.e. the rows of a chunk are contiguous - so a view shifted by
// s rows is just base + 16*s, and the second chunk of a K=16 step can live anywhere above the first (LBO is free).
// Checks, against a host reference, one tcgen05.mma (M=128, N=96, K=16):
//   variant 0: LBO in bits [16,30), SBO in bits [32,46)          (the documented order)
//   variant 1: the two fields swapped
// each with (a) an unshifted A view, (b) A shifted by 1 and 2 rows, (c) the second A chunk redirected to a zero region.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I deepdenoiser_b200/csrc -o /tmp/umma_nosw_probe tools/umma_nosw_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dd_ptx.cuh"

using namespace dd;

constexpr int kRowsA = 136;          // pixels per chunk plane of A
constexpr int kRowsB = 96;

__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int swap) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  const uint32_t f16 = (swap ? sbo_bytes : lbo_bytes) >> 4, f32 = (swap ? lbo_bytes : sbo_bytes) >> 4;
  d |= static_cast<uint64_t>(f16 & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(f32 & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;                                      // layout type 0 = no swizzle
}

struct Params {
  int swap;          // descriptor field order under test (one per process: a wrong order may fault)
  const __half* a;   // [2 chunks][kRowsA][8]
  const __half* b;   // [2 chunks][kRowsB][8]
  float* out;        // [cases][128][96]
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* a_s = smem;                            // 2 * 136 * 16 = 4352
  uint8_t* b_s = smem + 8192;                     // 2 * 96 * 16 = 3072
  uint8_t* z_s = smem + 16384;                    // zero region 4 KB
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * kRowsA * 8; i += 128) reinterpret_cast<__half*>(a_s)[i] = p.a[i];
  for (int i = tid; i < 2 * kRowsB * 8; i += 128) reinterpret_cast<__half*>(b_s)[i] = p.b[i];
  for (int i = tid; i < 4096 / 4; i += 128) reinterpret_cast<uint32_t*>(z_s)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_f16(128, 96);
  uint32_t phase = 0;
  int c = 0;
  {
    const int swap = p.swap;
    for (int kind = 0; kind < 4; ++kind, ++c) {
      // kind 0..2: A view shifted by `kind` rows, chunk stride = plane size; kind 3: second chunk -> zero region
      if (tid == 0) {
        const uint32_t a_addr = smem_u32(a_s) + (kind < 3 ? kind * 16 : 0);
        const uint32_t a_lbo = (kind < 3) ? kRowsA * 16 : (smem_u32(z_s) - smem_u32(a_s));
        const uint64_t ad = make_desc_nosw(a_addr, a_lbo, 128, swap);
        const uint64_t bd = make_desc_nosw(smem_u32(b_s), kRowsB * 16, 128, swap);
        umma_f16(tmem, ad, bd, idesc, 0u);
        umma_commit(&bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      for (int cb = 0; cb < 96; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + cb, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) p.out[(static_cast<size_t>(c) * 128 + tid) * 96 + cb + i] = __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncthreads();
    }
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

int main(int argc, char** argv) {
  const int swap_arg = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<__half> a(2 * kRowsA * 8), b(2 * kRowsB * 8);
  std::vector<float> af(a.size()), bf(b.size());
  srand(7);
  for (size_t i = 0; i < a.size(); ++i) { af[i] = (rand() % 17 - 8) / 8.f; a[i] = __float2half(af[i]); }
  for (size_t i = 0; i < b.size(); ++i) { bf[i] = (rand() % 13 - 6) / 4.f; b[i] = __float2half(bf[i]); }
  Params p;
  __half *da, *db; float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 8 * 128 * 96 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  p.a = da; p.b = db; p.out = dout; p.swap = swap_arg;
  // 200 KB: whatever the field order means, every address a descriptor can form stays inside the allocation
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe_kernel<<<1, 128, 200 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out(8 * 128 * 96);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  int c = 0;
  for (int swap = swap_arg; swap <= swap_arg; ++swap)
    for (int kind = 0; kind < 4; ++kind, ++c) {
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 96; ++n) {
          double want = 0;
          const int shift = kind < 3 ? kind : 0;
          for (int k = 0; k < 16; ++k) {
            const int j = k / 8, e8 = k % 8;
            const float av = (kind == 3 && j == 1) ? 0.f : af[(j * kRowsA + m + shift) * 8 + e8];
            want += static_cast<double>(av) * bf[(j * kRowsB + n) * 8 + e8];
          }
          const double d = fabs(out[(static_cast<size_t>(c) * 128 + m) * 96 + n] - want);
          if (d > worst) worst = d;
        }
      printf("variant %d (%s) kind %d (%s): max |err| = %.4g  %s\n", swap, swap ? "fields swapped" : "LBO@16 SBO@32", kind,
             kind < 3 ? "A shifted by `kind` rows" : "second A chunk -> zero region", worst, worst < 1e-3 ? "OK" : "MISMATCH");
    }
  return 0;
}

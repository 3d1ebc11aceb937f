// Microbenchmark: ceiling of the observation write path.  Persistent CTAs, each warp owns a
// shared-memory staging tile of ROWS x F floats, fills it (F scalar st.shared per lane, like the row
// assembly of the step kernel) and writes it to HBM either with one bulk (TMA) store or with a
// coalesced st.global.v4 loop.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store_bw tma_store_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int F = 51;
constexpr int ROWS = 32;
constexpr int CHUNK = ROWS * F;  // floats per warp tile (6528 B)

__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* g, const void* s, uint32_t bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(sa), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// mode 0: TMA, one buffer per warp; mode 1: TMA, two buffers per warp (wait_group.read 1); mode 2: st.global.v4
template <int MODE>
__global__ void __launch_bounds__(256) store_kernel(float* out, int n_chunks, int warps_per_cta, int fill) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps_per_cta) return;
  constexpr int NBUF = MODE == 1 ? 2 : 1;
  float* stage = smem + warp * NBUF * CHUNK;
  int it = 0;
  for (int c = blockIdx.x * warps_per_cta + warp; c < n_chunks; c += gridDim.x * warps_per_cta, ++it) {
    float* buf = stage + (NBUF == 2 ? (it & 1) * CHUNK : 0);
    if (MODE != 2) {
      if (lane == 0) bulk_wait_read<NBUF - 1>();
      __syncwarp();
    }
    if (fill) {
      float* row = buf + lane * F;
#pragma unroll
      for (int k = 0; k < F; ++k) row[k] = (float)(c + k);
    }
    float* dst = out + (size_t)c * CHUNK;
    if (MODE != 2) {
      fence_async();
      __syncwarp();
      if (lane == 0) bulk_store(dst, buf, CHUNK * 4);
    } else {
      __syncwarp();
      const float4* s4 = reinterpret_cast<const float4*>(buf);
      float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
      for (int i = lane; i < CHUNK / 4; i += 32) d4[i] = s4[i];
      __syncwarp();
    }
  }
  if (MODE != 2 && lane == 0) bulk_wait_read<0>();
}

template <int MODE>
static void run(const char* name, float* out, int n_chunks, int ctas_per_sm, int warps, int fill) {
  const int nbuf = MODE == 1 ? 2 : 1;
  const size_t smem = (size_t)warps * nbuf * CHUNK * 4;
  cudaFuncSetAttribute(store_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, store_kernel<MODE>, 256, smem);
  if (occ < ctas_per_sm) { printf("%-28s ctas/sm %d warps %d: occupancy %d too low\n", name, ctas_per_sm, warps, occ); return; }
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) store_kernel<MODE><<<grid, 256, smem>>>(out, n_chunks, warps, fill);
  cudaEventRecord(e0);
  const int reps = 50;
  for (int i = 0; i < reps; ++i) store_kernel<MODE><<<grid, 256, smem>>>(out, n_chunks, warps, fill);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / reps;
  const double bytes = (double)n_chunks * CHUNK * 4;
  printf("%-28s ctas/sm %d warps %d fill %d: %.1f us  %.0f GB/s  (%s)\n", name, ctas_per_sm, warps, fill, us, bytes / us / 1e3,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int n_chunks = 1638400 / ROWS;  // c4: 1,638,400 rows of 51 floats = 334 MB
  float* out;
  cudaMalloc(&out, (size_t)n_chunks * CHUNK * 4);
  for (int fill = 0; fill <= 1; ++fill) {
    for (int cps = 1; cps <= 4; ++cps) {
      run<0>("tma 1 buffer/warp", out, n_chunks, cps, 7, fill);
      run<1>("tma 2 buffers/warp", out, n_chunks, cps, 7, fill);
      run<2>("st.global.v4", out, n_chunks, cps, 7, fill);
    }
    run<0>("tma 1 buffer/warp", out, n_chunks, 4, 8, fill);
    run<0>("tma 1 buffer/warp", out, n_chunks, 2, 4, fill);
  }
  return 0;
}

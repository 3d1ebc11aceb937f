// Microbenchmark: what write pattern reaches the HBM write ceiling on B200?  334 MB (the c4 observation tensor) per
// launch, steady state (back-to-back launches).
//   fill       : torch-like grid-stride st.global.v4, fully coalesced, non-persistent, 1024 threads x many blocks
//   chunk_np   : one warp per 6528-byte chunk (32 rows x 51 floats), non-persistent grid (STG.v4 from registers)
//   chunk_p    : the same chunks, persistent CTAs (148 x k), strided chunk order (what the step kernel does)
//   chunk_pc   : persistent CTAs, each CTA walks a CONTIGUOUS range of chunks
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <type_traits>

constexpr int CHUNK4 = 32 * 51 / 4;  // float4 per chunk (408)

__global__ void fill_kernel(float4* out, size_t n4) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}

template <int ORDER>  // 0: chunk = global warp id, strided by total warps; 1: contiguous range per CTA
__global__ void __launch_bounds__(256) chunk_kernel(float4* out, int n_chunks, int warps_per_cta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps_per_cta) return;
  const float4 v = make_float4(1.f, 2.f, 3.f, (float)warp);
  if (ORDER == 0) {
    for (int c = blockIdx.x * warps_per_cta + warp; c < n_chunks; c += gridDim.x * warps_per_cta) {
      float4* d = out + (size_t)c * CHUNK4;
      for (int i = lane; i < CHUNK4; i += 32) d[i] = v;
    }
  } else {
    const int per_cta = (n_chunks + gridDim.x - 1) / gridDim.x;
    const int c0 = blockIdx.x * per_cta, c1 = min(n_chunks, c0 + per_cta);
    for (int c = c0 + warp; c < c1; c += warps_per_cta) {
      float4* d = out + (size_t)c * CHUNK4;
      for (int i = lane; i < CHUNK4; i += 32) d[i] = v;
    }
  }
}

// persistent CTAs, chunks claimed dynamically in global order: per warp (GRAN = 1) or 7 consecutive chunks per CTA
// (GRAN = 7, one atomic per CTA round).  `base` = launch index * n_chunks, the counter is never reset.
template <int GRAN>
__global__ void __launch_bounds__(256) chunk_dynamic_kernel(float4* out, int n_chunks, unsigned long long* counter, unsigned long long base) {
  __shared__ unsigned long long s_first;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 v = make_float4(1.f, 2.f, 3.f, (float)warp);
  if (GRAN == 1) {
    if (warp >= 7) return;
    for (;;) {
      unsigned long long c = 0;
      if (lane == 0) c = atomicAdd(counter, 1ull) - base;
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= (unsigned long long)n_chunks) break;
      float4* d = out + (size_t)c * CHUNK4;
      for (int i = lane; i < CHUNK4; i += 32) d[i] = v;
    }
  } else {
    for (;;) {
      __syncthreads();
      if (threadIdx.x == 0) s_first = atomicAdd(counter, (unsigned long long)GRAN) - base;
      __syncthreads();
      const unsigned long long first = s_first;
      if (first >= (unsigned long long)n_chunks) break;
      for (int r = 0; r < GRAN; r += 7) {  // the claimed tiles one after the other, 7 chunks (one tile) at a time
        const unsigned long long c = first + r + warp;
        if (warp < 7 && c < (unsigned long long)n_chunks) {
          float4* d = out + (size_t)c * CHUNK4;
          for (int i = lane; i < CHUNK4; i += 32) d[i] = v;
        }
      }
    }
  }
}

template <typename F>
static void timeit(const char* name, size_t bytes, F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) launch();
  cudaEventRecord(e0);
  const int reps = 50;
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / reps;
  printf("%-44s %.1f us  %.0f GB/s  (%s)\n", name, us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int n_chunks = 1638400 / 32;
  const size_t n4 = (size_t)n_chunks * CHUNK4, bytes = n4 * 16;
  float4* out;
  cudaMalloc(&out, bytes);
  char name[128];
  for (int blocks : {148 * 2, 148 * 8, 148 * 32, 148 * 128}) {
    snprintf(name, sizeof(name), "fill grid-stride, %d blocks x 1024", blocks);
    timeit(name, bytes, [&] { fill_kernel<<<blocks, 1024>>>(out, n4); });
  }
  for (int warps : {7, 8}) {
    snprintf(name, sizeof(name), "chunk non-persistent, %d warps/CTA", warps);
    timeit(name, bytes, [&] { chunk_kernel<0><<<(n_chunks + warps - 1) / warps, 256>>>(out, n_chunks, warps); });
  }
  for (int cps : {1, 3, 8}) {
    snprintf(name, sizeof(name), "chunk persistent strided, %d CTAs/SM x 7 warps", cps);
    timeit(name, bytes, [&] { chunk_kernel<0><<<148 * cps, 256>>>(out, n_chunks, 7); });
    snprintf(name, sizeof(name), "chunk persistent contiguous, %d CTAs/SM x 7", cps);
    timeit(name, bytes, [&] { chunk_kernel<1><<<148 * cps, 256>>>(out, n_chunks, 7); });
  }
  unsigned long long* counter;
  cudaMalloc(&counter, 8);
  cudaMemset(counter, 0, 8);
  // the counter is never reset: launch L starts at base = L * (claims per launch), known on the host without a sync
  unsigned long long base = 0;
  for (int cps : {1, 3, 8}) {
    snprintf(name, sizeof(name), "chunk persistent DYNAMIC per warp, %d CTAs/SM", cps);
    const unsigned long long per_launch = (unsigned long long)n_chunks + 148ull * cps * 7;  // every warp overshoots once
    timeit(name, bytes, [&] {
      chunk_dynamic_kernel<1><<<148 * cps, 256>>>(out, n_chunks, counter, base);
      base += per_launch;
    });
  }
  {
    const int cps = 3;
    auto run_gran = [&](auto tag, const char* label) {
      constexpr int G = decltype(tag)::value;
      snprintf(name, sizeof(name), "chunk persistent DYNAMIC per CTA (%s), %d CTAs/SM", label, cps);
      const unsigned long long per_launch = (unsigned long long)G * (((unsigned long long)n_chunks + G - 1) / G + 148ull * cps);
      timeit(name, bytes, [&] {
        chunk_dynamic_kernel<G><<<148 * cps, 256>>>(out, n_chunks, counter, base);
        base += per_launch;
      });
    };
    run_gran(std::integral_constant<int, 7>{}, "1 tile");
    run_gran(std::integral_constant<int, 14>{}, "2 tiles");
    run_gran(std::integral_constant<int, 28>{}, "4 tiles");
    run_gran(std::integral_constant<int, 56>{}, "8 tiles");
  }
  return 0;
}

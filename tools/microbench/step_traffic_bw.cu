// Microbenchmark: steady-state time of the step kernel's HBM traffic pattern with no compute and no
// cross-warp synchronisation.  Per warp and tile of 32 houses: read the per-house inputs, write the
// state + reward, write the 32 x 51-float observation tile with one bulk (TMA) store.
//   layout 0 (SoA, as shipped): coef_a 16 B, coef_b 16 B, temps 8 B, coef_c 8 B, hvac 4 B, action 1 B per house
//   layout 1 (AoS): one 64-byte record per house (same fields + padding), state written back into the record
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o step_traffic_bw step_traffic_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int F = 51, ROWS = 32, CHUNK = ROWS * F;

__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* g, const void* s, uint32_t bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(sa), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

struct Bufs {
  float4 *coef_a, *coef_b;
  float2 *temps, *coef_c;
  int* hvac;
  unsigned char* act;
  float* reward;
  float* obs;
  float4* rec;  // AoS: 4 x float4 per house
};

template <int LAYOUT, bool OBS>
__global__ void __launch_bounds__(256) traffic_kernel(Bufs b, int n_chunks, int warps_per_cta) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps_per_cta) return;
  float* stage = smem + warp * CHUNK;
  for (int c = blockIdx.x * warps_per_cta + warp; c < n_chunks; c += gridDim.x * warps_per_cta) {
    const size_t h = (size_t)c * 32 + lane;
    float acc;
    if (LAYOUT == 0) {
      const float4 a = b.coef_a[h], bb = b.coef_b[h];
      const float2 t = b.temps[h], cc = b.coef_c[h];
      const int hv = b.hvac[h];
      const int act = b.act[h];
      acc = a.x + a.w + bb.y + bb.z + t.x + t.y + cc.x + cc.y + (float)(hv + act);
      b.temps[h] = make_float2(acc, t.y + 1.0f);
      b.hvac[h] = hv + 1;
    } else {
      // coalesced: lane l copies 16-byte chunk i*32+l of the warp's 2 KB of records
      const float4* src = b.rec + (size_t)c * 128;
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = src[i * 32 + lane];
      const int act = b.act[h];
      acc = v[0].x + v[1].y + v[2].z + v[3].w + (float)act;
      // state lives in chunk 2 of each record (every 4th 16-byte chunk of the 2 KB): lanes holding one write it back
      if ((lane & 3) == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 o = v[i];
          o.x = acc; o.y += 1.0f; o.z += 1.0f;
          const_cast<float4*>(src)[i * 32 + lane] = o;
        }
      }
    }
    b.reward[h] = acc;
    if (OBS) {
      if (lane == 0) bulk_wait_read();
      __syncwarp();
      float* row = stage + lane * F;
#pragma unroll
      for (int k = 0; k < F; ++k) row[k] = acc + (float)k;
      fence_async();
      __syncwarp();
      if (lane == 0) bulk_store(b.obs + (size_t)c * CHUNK, stage, CHUNK * 4);
    }
  }
  if (OBS && lane == 0) bulk_wait_read();
}

template <int LAYOUT, bool OBS>
static void run(const char* name, Bufs b, int n_chunks, int ctas_per_sm) {
  const int warps = 7;
  const size_t smem = (size_t)warps * CHUNK * 4 + 34 * 1024;  // pad to the step kernel's footprint (3 CTAs/SM max)
  cudaFuncSetAttribute(traffic_kernel<LAYOUT, OBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) traffic_kernel<LAYOUT, OBS><<<grid, 256, smem>>>(b, n_chunks, warps);
  cudaEventRecord(e0);
  const int reps = 100;
  for (int i = 0; i < reps; ++i) traffic_kernel<LAYOUT, OBS><<<grid, 256, smem>>>(b, n_chunks, warps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-40s ctas/sm %d: %.1f us/step  (%s)\n", name, ctas_per_sm, ms * 1e3 / reps, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const size_t H = 1638400;  // c4: houses per GPU
  const int n_chunks = (int)(H / 32);
  Bufs b;
  cudaMalloc(&b.coef_a, H * 16); cudaMalloc(&b.coef_b, H * 16); cudaMalloc(&b.temps, H * 8); cudaMalloc(&b.coef_c, H * 8);
  cudaMalloc(&b.hvac, H * 4); cudaMalloc(&b.act, H); cudaMalloc(&b.reward, H * 4); cudaMalloc(&b.obs, H * F * 4);
  cudaMalloc(&b.rec, H * 64);
  cudaMemset(b.coef_a, 0, H * 16); cudaMemset(b.coef_b, 0, H * 16); cudaMemset(b.temps, 0, H * 8); cudaMemset(b.coef_c, 0, H * 8);
  cudaMemset(b.hvac, 0, H * 4); cudaMemset(b.act, 0, H); cudaMemset(b.rec, 0, H * 64);
  for (int cps = 1; cps <= 3; ++cps) {
    run<0, true>("SoA inputs + state + reward + obs", b, n_chunks, cps);
    run<1, true>("AoS 64-B records + reward + obs", b, n_chunks, cps);
    run<0, false>("SoA inputs + state + reward (no obs)", b, n_chunks, cps);
    run<1, false>("AoS 64-B records + reward (no obs)", b, n_chunks, cps);
  }
  return 0;
}

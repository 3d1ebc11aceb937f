// Host-side expansion of the compact observation records (csrc/mdr_expand.h) timed on this machine's cores:
//   g++ -O3 -std=c++17 -pthread -o expand_bw expand_bw.cpp && ./expand_bw [threads] [envs] [houses]
// Prints GB/s of observation bytes written and checks the rows against a scalar restatement.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../marl-demandresponse-original_b200/csrc/mdr_expand.h"

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 8, E = argc > 2 ? atoi(argv[2]) : 16384, N = argc > 3 ? atoi(argv[3]) : 100, C = 10;
  const int F = 11 + 4 * C;
  std::vector<float> compact((size_t)E * N * 16);
  for (size_t i = 0; i < compact.size(); ++i) compact[i] = (float)((i * 2654435761u) % 1000) * 0.001f + 0.5f;
  float* obs = nullptr;
  if (posix_memalign((void**)&obs, 64, (size_t)E * N * F * sizeof(float))) return 1;
  printf("stores: %s\n", mdr::use_stream_stores() ? "non-temporal" : "plain (MDR_HOST_NT=0)");
  for (int rep = 0; rep < 4; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int w = 0; w < T; ++w)
      th.emplace_back([&, w] {
        std::vector<float> block;
        mdr::expand_envs<float>(compact.data(), obs, (int)((long long)E * w / T), (int)((long long)E * (w + 1) / T), N, C, block);
      });
    for (auto& t : th) t.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %d: %.2f ms, %.1f GB/s written (%.1f GB/s per thread)\n", T, s * 1e3, (double)E * N * F * 4 / s / 1e9,
           (double)E * N * F * 4 / s / 1e9 / T);
  }
  // check against the definition
  size_t bad = 0;
  for (int e = 0; e < E; e += 997)
    for (int i = 0; i < N; ++i) {
      const float* c = compact.data() + (size_t)e * N * 16;
      const float* row = obs + ((size_t)e * N + i) * F;
      for (int k = 0; k < 11; ++k) bad += row[k] != c[i * 16 + k];
      for (int k = 0; k < C; ++k) {
        int j = k < C / 2 ? i - C / 2 + k : i + 1 + (k - C / 2);
        j = (j + N) % N;
        const float* m = c + j * 16 + 11;
        bad += row[11 + 4 * k] != m[0];
        bad += row[12 + 4 * k] != m[1] * c[i * 16 + 15];
        bad += row[13 + 4 * k] != m[2];
        bad += row[14 + 4 * k] != m[3];
      }
    }
  printf("mismatches: %zu\n", bad);
  return bad != 0;
}

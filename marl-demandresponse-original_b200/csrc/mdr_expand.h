// Host side of the compact observation transfer (mdr_compact.cuh, mdr_host.cu): expands the 16-real records of
// envs [e0, e1) into rows of F = 11 + 4*C reals in utils.normStateDict order (utils.py:774-868).  Pure C++ + SSE2
// (x86-64 baseline), no CUDA: tools/microbench/expand_bw.cpp times it on the build container's CPU.
//   row(i) = own[0..10] ++ for each of the C neighbours j of house i (:816-828: the C houses around i, skipping i,
//            wrapping around the env): (dT_j/5, sso_j * (1/lockout_i), P_j/7500, Pmax_j/7500)
// The single multiply is the same IEEE operation the step kernels' row assembly performs on the same operands, so the
// expanded rows are bit-identical to the rows a kernel writes.
#pragma once
#include <emmintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace mdr {

// streams `bytes` (multiple of 16) from src to the 16-byte aligned dst past the caches: the caller's buffer is written
// once and not read here, so no read-for-ownership traffic
static inline void stream_out(void* dst, const void* src, size_t bytes) {
  const __m128i* s = reinterpret_cast<const __m128i*>(src);
  __m128i* d = reinterpret_cast<__m128i*>(dst);
  size_t q = 0;
  const size_t n = bytes / 16;
  for (; q + 4 <= n; q += 4) {
    const __m128i a = _mm_loadu_si128(s + q), b = _mm_loadu_si128(s + q + 1), c = _mm_loadu_si128(s + q + 2),
                  e = _mm_loadu_si128(s + q + 3);
    _mm_stream_si128(d + q, a);
    _mm_stream_si128(d + q + 1, b);
    _mm_stream_si128(d + q + 2, c);
    _mm_stream_si128(d + q + 3, e);
  }
  for (; q < n; ++q) _mm_stream_si128(d + q, _mm_loadu_si128(s + q));
}

static inline void expand_row(const float* c, int i, int N, int C, float* b) {
  const int half = C >> 1;
  const float* own = c + (size_t)i * 16;
  _mm_storeu_ps(b, _mm_loadu_ps(own));
  _mm_storeu_ps(b + 4, _mm_loadu_ps(own + 4));
  _mm_storeu_ps(b + 7, _mm_loadu_ps(own + 7));  // features 7..10 (overlapping store: 11 is not a multiple of 4)
  const __m128 scale = _mm_set_ps(1.0f, 1.0f, own[15], 1.0f);  // (x, y * inv_lock, z, w); x * 1.0f == x exactly
  float* m = b + 11;
  int j = i - half;
  if (j < 0) j += N;
  for (int k = 0; k < C; ++k, m += 4) {
    if (k == half) { if (++j >= N) j -= N; }  // skip the house itself
    _mm_storeu_ps(m, _mm_mul_ps(_mm_loadu_ps(c + (size_t)j * 16 + 11), scale));
    if (++j >= N) j -= N;
  }
}

static inline void expand_row(const double* c, int i, int N, int C, double* b) {
  const int half = C >> 1;
  const double* own = c + (size_t)i * 16;
  for (int k = 0; k < 11; ++k) b[k] = own[k];
  const double inv_lock = own[15];
  double* m = b + 11;
  int j = i - half;
  if (j < 0) j += N;
  for (int k = 0; k < C; ++k, m += 4) {
    if (k == half) { if (++j >= N) j -= N; }
    const double* s = c + (size_t)j * 16 + 11;
    m[0] = s[0]; m[1] = s[1] * inv_lock; m[2] = s[2]; m[3] = s[3];
    if (++j >= N) j -= N;
  }
}

// how the expanded rows reach the caller's buffer: 1 = non-temporal stores (default), 0 = plain memcpy (write-allocate);
// MDR_HOST_NT=0 selects the latter (A/B on hosts where one of the two is clearly faster)
static inline bool use_stream_stores() {
  static const bool v = [] { const char* s = getenv("MDR_HOST_NT"); return !(s && s[0] == '0'); }();
  return v;
}

template <typename R>
static void expand_envs(const R* compact, R* obs, int e0, int e1, int N, int C, std::vector<R>& block) {
  const int F = 11 + 4 * C;
  const size_t env_elems = (size_t)N * F;
  // a block of whole envs of about 32 KB stays in L1/L2 between being assembled and being streamed out
  int per = (int)(32768 / (env_elems * sizeof(R)));
  if (per < 1) per = 1;
  block.resize(env_elems * per + 4);
  for (int e = e0; e < e1; e += per) {
    const int ne = e1 - e < per ? e1 - e : per;
    R* b = block.data();
    for (int q = 0; q < ne; ++q) {
      const R* c = compact + (size_t)(e + q) * N * 16;
      for (int i = 0; i < N; ++i, b += F) expand_row(c, i, N, C, b);
    }
    R* dst = obs + (size_t)e * env_elems;
    const size_t bytes = env_elems * ne * sizeof(R);
    if (use_stream_stores() && ((reinterpret_cast<uintptr_t>(dst) | bytes) & 15) == 0) stream_out(dst, block.data(), bytes);
    else memcpy(dst, block.data(), bytes);
  }
  _mm_sfence();
}

}  // namespace mdr

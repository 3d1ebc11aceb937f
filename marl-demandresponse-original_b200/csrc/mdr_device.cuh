// Device-side helpers of the fused MADemandResponseEnv step kernel (sm_100a only).
// Reference citations are relative to zhimaerfan/marl-demandresponse-original.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mdr_b200.h"

namespace mdr {

template <typename R> struct Vec;
template <> struct Vec<float> { using T2 = float2; using T4 = float4; };
template <> struct Vec<double> { using T2 = double2; using T4 = double4; };

__device__ __forceinline__ float2 make2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ double2 make2(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ float4 make4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
__device__ __forceinline__ double4 make4(double a, double b, double c, double d) { return make_double4(a, b, c, d); }

// 1 / lockout_duration as every kernel (and the compact host transfer) computes it: the same bits everywhere, so the
// observation a step kernel assembles and the one expanded on the host from the compact record are identical
__device__ __forceinline__ float inv_real(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ double inv_real(double x) { return 1.0 / x; }

// Non-contracted arithmetic: the interpolation and the grid signal are evaluated in the
// reference's operation order (scipy/_rgi.py:520-549, env/MA_DemandResponse.py:1295-1314) so
// that fp64 results agree to the last bit with numpy given identical inputs.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator: every (entity, step, stream) owns its own draw, so
// results do not depend on the launch geometry.
// ---------------------------------------------------------------------------------------
enum : uint32_t { STREAM_OD = 1, STREAM_ACT = 2, STREAM_IDS = 3, STREAM_MSG = 4, STREAM_PERLIN = 5 };

__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t key) {
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {  // (0, 1), 53 bits
  const uint64_t v = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);
  return ((double)(v & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

// standard normal from one Philox draw (production-mode outdoor-temperature noise; single
// precision Box-Muller is ample for a 0.5 K noise term and keeps the per-env prologue short)
__device__ __forceinline__ double normal_from(const uint4 r) {
  const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return (double)(sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2));
}

// ---------------------------------------------------------------------------------------
// Calendar (proleptic Gregorian, naive datetime like the reference's `datetime` objects)
// ---------------------------------------------------------------------------------------
struct Calendar {
  int year, month, day, yday;  // yday = tm_yday (1-based); only valid after calendar_date()
  int hour, minute, second, sod;
  int days;                    // whole days since 1970-01-01
};

// All 32-bit: t is the naive epoch second as uint32 (valid until 2106; validated on the host),
// so every division is by a compile-time constant (mul-shift), never a 64-bit software divide.
__device__ __forceinline__ Calendar calendar_time(uint32_t t) {
  Calendar c;
  const uint32_t days = t / 86400u;
  const uint32_t sod = t - days * 86400u;
  c.days = (int)days;
  c.sod = (int)sod;
  c.hour = (int)(sod / 3600u);
  c.minute = (int)((sod - (uint32_t)c.hour * 3600u) / 60u);
  c.second = (int)(sod - (uint32_t)c.hour * 3600u - (uint32_t)c.minute * 60u);
  c.year = c.month = c.day = c.yday = 0;
  return c;
}

__device__ __forceinline__ int days_from_civil(int y, int m, int d) {
  y -= m <= 2;
  const int era = y / 400;  // y >= 1970 here
  const int yoe = y - era * 400;
  const int doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const int doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}

// fills year / month / day / yday (needed only for solar gain, the day-of-year features and the
// interpolation point when solar gain is on)
__device__ __forceinline__ void calendar_date(Calendar& c) {
  const int z = c.days + 719468;
  const int era = z / 146097;
  const int doe = z - era * 146097;
  const int yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  const int doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const int mp = (5 * doy + 2) / 153;
  c.day = doy - (153 * mp + 2) / 5 + 1;
  c.month = mp < 10 ? mp + 3 : mp - 9;
  c.year = yoe + era * 400 + (c.month <= 2);
  c.yday = c.days - days_from_civil(c.year, 1, 1) + 1;
}

// utils.py:1277-1350 -- CIBSE solar cooling load polynomial, same term order.
__device__ __forceinline__ double solar_gain(const Calendar& c, double window_area, double shading_coeff) {
  // c must have been completed with calendar_date()
  const double x = c.hour + c.minute / 60.0 - 7.5;
  double scl = 0.0;
  if (!(x < 0 || x > 10)) {
    const double y = c.month + c.day / 30.0 - 1;
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2, y2 = y * y, y3 = y2 * y, y4 = y2 * y2;
    scl = 4.36579418e01 + x * 1.58055357e02 + y * 8.76635241e01 + x2 * -4.55944821e01 +
          x2 * y * 3.24275366e00 + x2 * y2 * -4.56096472e-01 + y2 * -1.47795612e01 +
          x * y2 * 4.68950855e00 + x * y * -3.73313090e01 + x3 * 5.78827663e00 + y3 * 1.04354810e00 +
          x3 * y * 2.12969604e-02 + x3 * y2 * 2.58881400e-03 + x3 * y3 * -5.11397219e-04 +
          x2 * y3 * 1.56398008e-02 + x * y3 * -1.18302764e-01 + x4 * -2.71446436e-01 + y4 * -3.97855577e-02;
  }
  return window_area * shading_coeff * scl;
}

// ---------------------------------------------------------------------------------------
// 1-D perlin (production mode only; parity mode replays the host value).  Same construction
// as utils.Perlin (utils.py:1231-1253) over perlin_noise.PerlinNoise: per octave o,
//   n_o(x) = sum_{i in {floor(xo), floor(xo)+1}} fade(1-|xo-i|) * g_o(i) * (xo-i),  xo = x*octaves
// with lattice gradients g_o(i) ~ U(-1,1) drawn from Philox instead of python's `random`.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double perlin_fade(double t) { return t * t * t * (t * (t * 6.0 - 15.0) + 10.0); }

__device__ __forceinline__ float perlin_fade_f(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }

// Lattice gradient g_o(i) in (-1, 1): a pure function of (lattice point, octave, env key).  A two-round
// multiply-xorshift integer hash ("lowbias32" constants) instead of a Philox block: the prologue warp evaluates
// 2 * nb_octaves of these per env and step on its critical path, and a noise texture needs decorrelation, not a
// cryptographic generator (the Philox streams stay for the draws that model random variables).
__device__ __forceinline__ float perlin_gradient(int lattice, int octave, uint64_t key) {
  uint32_t x = ((uint32_t)lattice * 0x9E3779B1u) ^ (uint32_t)key ^ (((uint32_t)(key >> 32)) + (uint32_t)octave * 0x85EBCA77u);
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return 2.0f * (((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f)) - 1.0f;
}

// One octave of utils.Perlin.calculate_noise (utils.py:1247-1253) at position x = t / period: both lattice corners
// of PerlinNoise(octaves = 2^j * octaves_step), weighted 1 / 2^j (last octave: 1 / (2^nb - 1)).  One fp64 multiply and
// floor per octave, the rest in fp32 (the value is a noise texture; the per-octave result is the unit both the
// pipelined/generic prologue and the fused kernel's record warp sum in fp64, so they agree to the last bit or two).
__device__ __forceinline__ float perlin_octave(double x, int j, int nb, int octaves_step, uint64_t key) {
  const double xo = x * (double)((1 << j) * octaves_step);
  const double fl = floor(xo);
  const float d0 = (float)(xo - fl);  // distance to the left corner, [0, 1); to the right corner: d0 - 1
  const int i0 = (int)fl;
  const float v = perlin_fade_f(1.0f - d0) * perlin_gradient(i0, j, key) * d0 +
                  perlin_fade_f(d0) * perlin_gradient(i0 + 1, j, key) * (d0 - 1.0f);
  const float wgt = j == nb - 1 ? 1.0f / (float)((1 << nb) - 1) : 1.0f / (float)(1 << j);
  return v * wgt;
}

// ---------------------------------------------------------------------------------------
// warp reductions (fixed shuffle tree => run-to-run deterministic)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------
// bulk (TMA) shared -> global store of a contiguous, 16-byte aligned tile (SASS: UBLKCP)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

}  // namespace mdr

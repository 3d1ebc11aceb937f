// MADemandResponseEnv step path for B200 (sm_100a): one CTA owns G whole envs (clusters), one thread owns
// one house.  Replaces the Python object loops of
//   ClusterHouses.step        env/MA_DemandResponse.py:1005-1055
//   HVAC.step                 env/MA_DemandResponse.py:463-492
//   SingleHouse.update_temperature  :664-738
//   compute_rewards           :330-373
//   PowerGrid.step            :1236-1316 (+ interpolatePower :1195-1234)
//   make_cluster_obs_dict + normStateDict  :904-1003, utils.py:740-880
// This file: shared device code (interpolation, grid signal, per-env prologue), the reset-time precompute
// kernel, the generic step kernel (every mode, fp32/fp64, reset/observe, one tile per CTA) and the launchers.
//   mdr_pipe.cuh      persistent software-pipelined fp32 kernel of the default configuration (the hot kernel)
//   mdr_fused.cuh     K steps per launch with the house state in registers + metric accumulators (deploy loop)
//   mdr_populate.cuh  device-side population draw / masked partial reset
// Phases of the generic kernel (block barriers between them):
//   0  per env   : a dedicated warp advances the clock, draws/replays the outdoor temperature and noise,
//                  evaluates the grid signal (concurrently with phase A)
//   A  per house : load packed state/coefficients (8/16-byte coalesced), [greedy controller], lockout state
//                  machine, 2x2 affine ETP update, store state, stage message + power in shared memory
//   B  per env   : cluster power from warp partials; mean/max penalties for the common_* reward modes
//   C/D (refresh steps only) per house multilinear table interpolation -> per env base power
//   E  per house : reward, observation row assembled in a per-warp shared-memory tile from the
//                  neighbours' staged messages, tile written with one bulk (TMA) store
#include "mdr_kernels.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "mdr_device.cuh"

#ifndef MDR_BLOCKS_256
#define MDR_BLOCKS_256 3  // resident 256-thread CTAs per SM the compiler must allow (register cap 85)
#endif

namespace mdr {


// Step index of the Philox streams: the host's counter plus an optional device-resident one (CUDA-graph replays of a
// captured rollout must not repeat their draws: the graph advances the device counter, the baked-in host value stays).
__device__ __forceinline__ uint64_t step_now(const KernelParams& p) {
  return p.step_counter != nullptr ? p.step_index + *p.step_counter : p.step_index;
}

// ----------------------------------------------------------------------------------------
struct EnvScratch {
  double P, od_new, rew_sig, pen_mean, pen_max, hour_s, date, s_old, sig_new, sig_noise, base;
  double f_sig, f_pow, f_od, f_sin_day, f_cos_day, f_sin_hr, f_cos_hr, f_solar, gain_now;
  uint32_t t_new;
  int due, tsi, time_sec;
};

// Compact hand-over record of the pipelined kernel (prologue warp -> house warps), one per env
// and ring slot.  64 bytes; the pipelined kernel never runs with solar gain or calendar features,
// so the interpolation point's hour/date are 0 and need no slot here.
struct PipeEnv {
  double s_old;      // grid signal before this step (reward)
  double od_new;     // outdoor temperature after this step (interpolation point)
  double sig_noise;  // perlin value of this step (signal re-evaluated after a refresh)
  float od_old;      // outdoor temperature the thermal update uses
  float f_sig;       // normalised new signal = observation feature 9 (rewritten by a refresh)
  int due;           // interpolation refresh due for this env
  int time_sec;
  double sig_new;    // grid signal after this step (metric accumulators; not final when `due`)
  float gain;        // solar gain of this step (0 with solar gain off), utils.py:1277-1350
  uint32_t t_new;    // clock after this step (naive epoch seconds)
  double base;       // base power the signal was evaluated with (not final when `due`)
};
static_assert(sizeof(PipeEnv) == 64, "PipeEnv must stay 64 bytes");

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// shared-memory carve-up (must match between host sizing and the kernel)
struct SmemLayout {
  size_t off_msg, off_pw, off_val, off_pen, off_env, off_met, off_cl, off_stage, total;
};

// hmax = threads per CTA (>= G*N); the message window holds G*(N+C) entries (wrap-around halos)
// part_slots: warp-partial slots per env (part_stride, or -- env split over a cluster -- the warps of one CTA)
inline SmemLayout smem_layout(int real_bytes, int hmax, int genvs_max, int nwarps, int rows_per_pass, int n_features,
                              bool need_val, bool need_pen, bool has_obs, int n_comm, int part_stride, bool need_met,
                              int part_slots) {
  SmemLayout L;
  size_t o = 0;
  if (part_slots <= 0) part_slots = part_stride;
  L.off_msg = o; o += align16((size_t)(hmax + genvs_max * n_comm) * 4 * real_bytes);
  L.off_pw = o;  o += align16((size_t)genvs_max * part_slots * sizeof(double));
  L.off_val = o; o += need_val ? align16((size_t)hmax * 3 * sizeof(double)) : 0;  // interpolation values / greedy sort scratch
  L.off_pen = o; o += need_pen ? align16((size_t)hmax * sizeof(double)) : 0;
  L.off_env = o; o += align16((size_t)genvs_max * sizeof(EnvScratch));
  L.off_met = o; o += need_met ? align16((size_t)genvs_max * part_slots * 5 * sizeof(double)) : 0;
  L.off_cl = o;  o += 128;  // ClusterTot
  L.off_stage = o;
  o += has_obs ? align16((size_t)nwarps * rows_per_pass * n_features * real_bytes) : 0;
  L.total = o;
  return L;
}

// ----------------------------------------------------------------------------------------
// multilinear interpolation, monteCarlo/interpolation.py:113-142 + scipy linear interpn.
// `key` = (flat index over the 4 nearest thermal-ratio cells) * 16 + nearest HVAC-power index.
// ----------------------------------------------------------------------------------------
// The interpolation grid (dims + axes) as the kernels read it: KernelParams holds it (constant bank when accessed from
// the kernel body), and the pipelined kernels keep a copy in shared memory -- their refresh code lives in out-of-line
// functions that only see KernelParams through a reference, where every access is a generic global load (measured:
// 14 us per refreshed tile, all of it dependent parameter loads).
struct InterpGrid {
  int interp_dims[MDR_INTERP_DIMS];
  double interp_axes[MDR_INTERP_DIMS][MDR_INTERP_MAX_AXIS];
};

template <typename G>
__device__ __forceinline__ double clip_axis(const G& p, int d, double v) {
  const double lo = p.interp_axes[d][0], hi = p.interp_axes[d][p.interp_dims[d] - 1];
  if (v > hi) v = hi;
  else if (v < lo) v = lo;
  return v;
}

template <typename R, typename G>
__device__ __forceinline__ double interp_eval_grid_body(const G& p, const void* table_v, int key, double air, double mass,
                                                        double od, double hour, double date) {
  const int dims[5] = {4, 5, 6, 8, 9};
  const double x[5] = {clip_axis(p, 4, air), clip_axis(p, 5, mass), clip_axis(p, 6, od), clip_axis(p, 8, hour),
                       clip_axis(p, 9, date)};
  int idx[5];
  double w[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int d = dims[k];
    const int n = p.interp_dims[d];
    // searchsorted(right) - 1.  Fixed trip count: every axis entry is then an immediate constant-bank operand of the
    // compare (a run-time trip count makes each one a dependent LDC: measured ~13 us per refreshed tile)
    int i = 0;
#pragma unroll
    for (int j = 1; j < MDR_INTERP_MAX_AXIS; ++j) i += (j < n && p.interp_axes[d][j] <= x[k]) ? 1 : 0;
    i = min(i, n - 2);
    idx[k] = i;
    w[k] = (x[k] - p.interp_axes[d][i]) / (p.interp_axes[d][i + 1] - p.interp_axes[d][i]);
  }
  const int therm = key >> 4, ih = key & 15;
  const R* __restrict__ table = reinterpret_cast<const R*>(table_v);
  const int n4 = p.interp_dims[4], n5 = p.interp_dims[5], n6 = p.interp_dims[6], n7 = p.interp_dims[7],
            n8 = p.interp_dims[8], n9 = p.interp_dims[9];
  // All 32 corner values are fetched first (independent loads in flight together: the table lives in L2 and a corner
  // after corner walk pays the L2 latency 32 times -- measured 11 us per refresh), then accumulated in scipy's order.
  const size_t s9 = 1, s8 = (size_t)n9, s7 = s8 * n8, s6 = s7 * n7, s5 = s6 * n6, s4 = s5 * n5;
  const size_t base_off = (((((size_t)therm * n4 + idx[0]) * n5 + idx[1]) * n6 + idx[2]) * n7 + ih) * s7 + (size_t)idx[3] * s8 +
                          (size_t)idx[4] * s9;
  R corner[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) {  // itertools.product order: first dimension slowest
    const size_t off = base_off + ((c >> 4) & 1) * s4 + ((c >> 3) & 1) * s5 + ((c >> 2) & 1) * s6 + ((c >> 1) & 1) * s8 + (c & 1) * s9;
    corner[c] = __ldg(table + off);
  }
  double value = 0.0;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const int b0 = (c >> 4) & 1, b1 = (c >> 3) & 1, b2 = (c >> 2) & 1, b3 = (c >> 1) & 1, b4 = c & 1;
    double weight = 1.0;
    weight = mul_rn(weight, b0 ? w[0] : 1.0 - w[0]);
    weight = mul_rn(weight, b1 ? w[1] : 1.0 - w[1]);
    weight = mul_rn(weight, b2 ? w[2] : 1.0 - w[2]);
    weight = mul_rn(weight, b3 ? w[3] : 1.0 - w[3]);
    weight = mul_rn(weight, b4 ? w[4] : 1.0 - w[4]);
    value = add_rn(value, mul_rn((double)corner[c], weight));
  }
  return value;
}

// Out of line (the step kernels' hot loops must not carry its 40-odd registers; inlining the body into the pipelined
// kernel's refresh was tried and spilled twice as much under that kernel's 80-register cap).
template <typename R, typename G>
__device__ __noinline__ double interp_eval_grid(const G& p, const void* table_v, int key, double air, double mass, double od,
                                                double hour, double date) {
  return interp_eval_grid_body<R, G>(p, table_v, key, air, mass, od, hour, date);
}

template <typename R>
__device__ __forceinline__ double interp_eval(const KernelParams& p, int key, double air, double mass, double od, double hour,
                                              double date) {
  return interp_eval_grid<R, KernelParams>(p, p.interp_table, key, air, mass, od, hour, date);
}

// PowerGrid.step signal shapes, env/MA_DemandResponse.py:1257-1314
__device__ __forceinline__ double grid_signal(const KernelParams& p, double base, int time_sec, double noise,
                                              double ratio, double max_power) {
  double sig;
  const double two_pi = 2.0 * 3.141592653589793;
  if (p.signal_mode == MDR_SIG_FLAT) {
    sig = base;
  } else if (p.signal_mode == MDR_SIG_SINUSOIDALS) {
    sig = base;
    for (int i = 0; i < p.n_sinusoids; ++i) {
      const double amp = mul_rn(base, p.sin_ratios[i]);
      sig = add_rn(sig, mul_rn(amp, sin(mul_rn(two_pi, (double)time_sec) / p.sin_periods[i])));
    }
  } else if (p.signal_mode == MDR_SIG_REGULAR_STEPS) {
    const double amplitude = p.steps_amplitude_per_hvac * p.N;
    const double r = base / amplitude;
    const double arg = fmod((double)time_sec, p.steps_period) - mul_rn(1.0 - r, p.steps_period);
    sig = amplitude * (arg >= 0.0 ? 1.0 : 0.0);  // np.heaviside(arg, 1)
  } else {                                         // perlin family
    sig = fmax(0.0, add_rn(base, mul_rn(mul_rn(base, p.perlin_amplitude), noise)));
  }
  sig = mul_rn(sig, ratio);
  return fmin(sig, max_power);
}

// ----------------------------------------------------------------------------------------
// precompute: per-house derived coefficients (reset-time, fp64 math)
// ----------------------------------------------------------------------------------------
template <typename R>
__global__ void precompute_kernel(const __grid_constant__ KernelParams p) {
  using T2 = typename Vec<R>::T2;
  using T4 = typename Vec<R>::T4;
  const size_t h = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= (size_t)p.E * p.N) return;
  const double ua = p.ua[h], cm = p.cm[h], ca = p.ca[h], hm = p.hm[h], cap = p.cap[h];
  const double dt = (double)p.dt;
  // env/MA_DemandResponse.py:704-711
  const double a = cm * ca / hm;
  const double b = cm * (ua + hm) / hm + ca;
  const double c = ua;
  const double disc = sqrt(b * b - 4.0 * a * c);
  const double r1 = (-b + disc) / (2.0 * a);
  const double r2 = (-b - disc) / (2.0 * a);
  // :722-723
  const double A3 = r1 * ca / hm + (ua + hm) / hm;
  const double A4 = r2 * ca / hm + (ua + hm) / hm;
  // x = T_air - T_ss, y = T_mass - T_ss with T_ss = T_od + Q_a/Ua (= d/c, :707) turns :713-735 into
  //   x' = x*e2 + A1*(e1-e2),  y' = x*A4*e2 + A1*(A3*e1 - A4*e2),  A1 = (s*x - (Hm/Ca)*y)/(r2-r1)
  const double em1 = expm1(r1 * dt), em2 = expm1(r2 * dt);
  const double e1 = em1 + 1.0, e2 = em2 + 1.0;
  const double s = r2 + (ua + hm) / ca;
  const double k = (em1 - em2) / (r2 - r1);
  const double kk = (A3 * e1 - A4 * e2) / (r2 - r1);
  const double d11 = em2 + k * s;
  const double m12 = -k * hm / ca;
  const double m21 = A4 * e2 + kk * s;
  const double d22 = -kk * hm / ca - 1.0;
  reinterpret_cast<T4*>(p.coef_a)[h] = make4((R)d11, (R)m12, (R)m21, (R)d22);
  // HVAC.get_Q :505, HVAC.max_consumption :436
  const double q_on = -1.0 * cap / (1.0 + p.hvac_latent);
  const double p_on = cap / p.hvac_cop;
  reinterpret_cast<T4*>(p.coef_b)[h] = make4((R)(1.0 / ua), (R)q_on, (R)p_on, (R)p.target[h]);
  reinterpret_cast<T2*>(p.coef_c)[h] = make2((R)p.deadband[h], (R)p.lockout_dur[h]);
  // nearest cell of the interpolation table on the 4 thermal ratios and the HVAC power,
  // monteCarlo/interpolation.py:120-133 after utils.clipInterpolationPoint
  if (p.interp_key != nullptr) {
    const double v[4] = {ua / p.def_ua, cm / p.def_cm, ca / p.def_ca, hm / p.def_hm};
    int key = 0;
    for (int d = 0; d < 4; ++d) {
      const double x = clip_axis(p, d, v[d]);
      int best = 0;
      double bd = fabs(p.interp_axes[d][0] - x);
      for (int j = 1; j < p.interp_dims[d]; ++j) {
        const double dist = fabs(p.interp_axes[d][j] - x);
        if (dist < bd) { bd = dist; best = j; }
      }
      key = key * p.interp_dims[d] + best;
    }
    const double x = clip_axis(p, 7, cap);
    int best = 0;
    double bd = fabs(p.interp_axes[7][0] - x);
    for (int j = 1; j < p.interp_dims[7]; ++j) {
      const double dist = fabs(p.interp_axes[7][j] - x);
      if (dist < bd) { bd = dist; best = j; }
    }
    p.interp_key[h] = key * 16 + best;
  }
}

// ----------------------------------------------------------------------------------------
// per-env prologue: everything of the step that does not depend on the houses' new state
// (clock, outdoor temperature, noise draws, solar gain, grid signal when no refresh is due).
// One THREAD per env, executed by a dedicated warp of the CTA concurrently with the house
// warps' global loads and thermal update, so its latency (fp64 sin, Philox) is off the
// critical path.
// ----------------------------------------------------------------------------------------
// (trace builds: timeline of the pipelined kernel's per-env prologue, row 22 of the trace buffer; see mdr_pipe.cuh)
#ifdef MDR_TRACE
__device__ void trace_stamp_fn(int w, int t, int k);
#define MDR_PRO_STAMP(k) do { if (kPipe) trace_stamp_fn(threadIdx.x >> 5, 22, k); } while (0)
#else
#define MDR_PRO_STAMP(k) do { } while (0)
#endif

template <bool kPipe>
__device__ __forceinline__ int env_prologue(const KernelParams& p, EnvScratch& es, PipeEnv& pe, int e2, int sub, int L,
                                            bool valid, bool reset, bool observe_only) {
  // `L` lanes (a power of two) cooperate on one env: the Philox draws of the production mode are
  // spread over them; everything else is computed redundantly and written by sub-lane 0.
  // all per-env loads up front (independent, so their latencies overlap)
  const uint32_t t = (uint32_t)p.t_epoch[e2] + (reset ? 0u : (uint32_t)p.dt);
  const double od_prev = p.od_temp[e2];
  const double s_old = p.signal[e2];
  const double phase = p.phase[e2];
  const double ratio = p.artificial_ratio[e2], max_power = p.max_power[e2];
  const bool perlin = p.signal_mode == MDR_SIG_PERLIN;
  const bool draw_od = !reset && p.od_noise == nullptr;
  const bool draw_perlin = perlin && !observe_only && p.signal_noise == nullptr;
  double od_noise = (!reset && !draw_od) ? p.od_noise[e2] : 0.0;
  double sig_noise = (perlin && !draw_perlin && !observe_only) ? p.signal_noise[e2] : 0.0;
  const bool interp_mode = p.base_power_mode == MDR_BASE_INTERPOLATION;
  double base = p.avg_power_per_hvac * p.N;  // PowerGrid.step :1248-1249
  int tsi = 0;
  if (interp_mode) {
    base = p.base_power[e2];
    tsi = p.time_since_interp[e2];
  }
  const double solar_prev = (observe_only && p.solar) ? p.solar_gain[e2] : 0.0;

  MDR_PRO_STAMP(1);
  Calendar cal = calendar_time(t);
  if (p.solar || (p.state_flags & MDR_STATE_DAY)) calendar_date(cal);
  MDR_PRO_STAMP(2);
  if (draw_od || draw_perlin) {  // warp-uniform
    // utils.Perlin.calculate_noise (utils.py:1247-1253) with hashed lattice gradients: item d < nb is octave d
    // (both lattice corners, see perlin_octave); the last item is the outdoor-temperature normal
    const int nb = p.perlin_nb_octaves;
    const int n_perlin = draw_perlin ? nb : 0;
    const int n_draws = n_perlin + (draw_od ? 1 : 0);
    const double x = (double)cal.sod * p.inv_perlin_period;  // time.mktime(...) % 86400 with TZ=UTC, :1297
    const uint64_t pkey = draw_perlin ? p.seed ^ (uint64_t)__double_as_longlong(p.perlin_seed[e2]) : 0;
    double terms = 0.0, normal = 0.0;
    for (int d = sub; d < n_draws; d += L) {
      if (d < n_perlin) {
        terms += (double)perlin_octave(x, d, nb, p.perlin_octaves_step, pkey);
      } else {
        normal = normal_from(philox4x32((uint32_t)(e2 + p.env_base), (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32), STREAM_OD,
                                        p.seed));
      }
    }
    for (int o = L >> 1; o > 0; o >>= 1) {
      terms += __shfl_xor_sync(0xffffffffu, terms, o);
      normal += __shfl_xor_sync(0xffffffffu, normal, o);
    }
    if (draw_perlin) sig_noise = terms;
    if (draw_od) od_noise = p.temp_std * normal;
  }
  MDR_PRO_STAMP(3);
  double od_new = od_prev;
  if (!reset) {
    // ClusterHouses.compute_OD_temp, :1070-1081 (amplitude, bias and 2*pi/24 folded on the host)
    const double time_day = cal.hour + cal.minute * (1.0 / 60.0);
    // (the fp32 pipelined kernel takes sinpi: no Payne-Hanek slow path in its instruction stream, and
    //  the 1e-16 relative difference to sin(2*pi/24 * u) is far below fp32 resolution)
    od_new = kPipe ? p.od_amplitude * sinpi((time_day + (-6 + phase)) * (1.0 / 12.0)) + p.od_bias
                   : p.od_amplitude * sin(p.two_pi_over_24 * (time_day + (-6 + phase))) + p.od_bias;
    od_new += od_noise;
  }
  // SingleHouse.update_temperature evaluates house_solar_gain at the NEW datetime (:694)
  double gain = (p.solar && !reset) ? solar_gain(cal, p.window_area, p.shading_coeff) : 0.0;
  if (observe_only) gain = solar_prev;
  const int time_sec = cal.hour * 3600 + cal.minute * 60 + cal.second;
  // PowerGrid.step, :1250-1255: the signal is final now unless an interpolation refresh is due
  int due = 0;
  if (interp_mode && !observe_only) {
    tsi += p.dt;
    due = tsi >= p.interp_update_period;
    if (due) tsi = 0;
  }
  MDR_PRO_STAMP(4);
  double sig = s_old;
  if (!observe_only && !due) sig = grid_signal(p, base, time_sec, sig_noise, ratio, max_power);
  MDR_PRO_STAMP(5);
  if (kPipe) {
    // pipelined kernel: the prologue warp owns the per-env outputs that neither depend on the houses
    // nor are read by them (each env belongs to exactly one tile per launch, so running ahead is safe)
    if (sub == 0 && valid) {
      pe.s_old = s_old;
      pe.od_new = od_new;
      pe.sig_noise = sig_noise;
      pe.od_old = (float)od_prev;
      pe.f_sig = (float)(sig * p.inv_norm_sig_agents);
      pe.due = due;
      pe.time_sec = time_sec;
      pe.sig_new = sig;
      pe.gain = (float)gain;
      pe.t_new = t;
      pe.base = base;
      // An env split over a cluster (mdr_pipe_split.cuh) is evaluated by the prologue warp of EVERY CTA of the cluster,
      // each for itself: nobody may overwrite the per-env state while a peer can still read it.  There the house
      // thread of rank 0 writes these outputs from the record, after the cluster's rendezvous of the tile.
      if (p.cl <= 1) {
        if (p.solar) p.solar_gain[e2] = gain;
        p.t_epoch[e2] = (int64_t)t;
        p.od_temp[e2] = od_new;  // the house threads take the OLD value from the record
        if (!due) {
          p.base_power[e2] = base;
          p.signal[e2] = sig;
          if (interp_mode) p.time_since_interp[e2] = tsi;
        } else {
          // deferred refresh (pipe_refresh_pass, after the tile loop of the same launch): the env is
          // marked with time_since_interp = -1 and its perlin value parked in base_power, both of
          // which the refresh overwrites
          p.time_since_interp[e2] = -1;
          p.base_power[e2] = sig_noise;
        }
      }
    }
    return valid ? due : 0;
  }
  if (sub == 0 && valid) {
    es.t_new = t;
    es.od_new = od_new;
    es.f_od = (od_new - 20) * 0.2;
    es.sig_noise = sig_noise;
    es.gain_now = gain;
    es.f_solar = gain * 1e-3;
    es.hour_s = p.solar ? (double)cal.sod : 0.0;  // interpolatePower point, :1198-1207
    es.date = p.solar ? (double)cal.yday : 0.0;
    es.time_sec = time_sec;
    if (p.state_flags & MDR_STATE_DAY) {
      es.f_sin_day = sin(cal.yday * 2 * 3.141592653589793 / 365);
      es.f_cos_day = cos(cal.yday * 2 * 3.141592653589793 / 365);
    }
    if (p.state_flags & MDR_STATE_HOUR) {
      es.f_sin_hr = sin(cal.hour * 2 * 3.141592653589793 / 24);
      es.f_cos_hr = cos(cal.hour * 2 * 3.141592653589793 / 24);
    }
    es.tsi = tsi;
    es.due = due;
    es.s_old = s_old;
    es.base = base;
    es.sig_new = sig;
    es.f_sig = sig * p.inv_norm_sig_agents;
  }
  return valid ? due : 0;
}

// CTA barriers as PTX: the dedicated prologue warp and the house warps arrive at barrier 0 from
// different program points (warp-uniform control flow, equal arrival counts)
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ void house_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ int cta_or(int pred) {
  int r;
  asm volatile("{ .reg .pred p, q; setp.ne.s32 p, %1, 0; bar.red.or.pred q, 0, p; selp.s32 %0, 1, 0, q; }"
               : "=r"(r)
               : "r"(pred)
               : "memory");
  return r;
}

// Thread-block-cluster rendezvous (one env split over the CTAs of a cluster, N > 224): every thread of every CTA of
// the cluster arrives and waits; release/acquire at cluster scope, so shared AND global writes made before it are
// visible to the peers after it.
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// generic address of `local` (a shared-memory address of this CTA) in the CTA of rank `rank` of the cluster (DSMEM)
template <typename T>
__device__ __forceinline__ const T* dsmem_peer(const T* local, int rank) {
  return reinterpret_cast<const T*>(__cluster_map_shared_rank(const_cast<T*>(local), (unsigned)rank));
}
// per-CTA totals exchanged through DSMEM when an env is split over a cluster
struct ClusterTot {
  double P, pen_sum, pen_max, pad;
  double met[8];  // metric partials (reward, offset, |error|, error^2 sums; max |error|)
};

// first env of this CTA's tile and whether the env is split over a cluster
__device__ __forceinline__ int tile_env0(const KernelParams& p) {
  return p.cl > 1 ? (int)(blockIdx.x / (unsigned)p.cl) : (int)blockIdx.x * p.G;
}

// Body of the dedicated prologue warp (and, when the CTA has no spare warp, of warp 0 before it
// turns to its houses).  Kept out of line so that none of its register pressure or call-saved
// state leaks into the house warps' code path.
__device__ __noinline__ int prologue_warp_main(const KernelParams& p, bool reset, bool observe_only, bool dedicated) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EnvScratch* s_env = reinterpret_cast<EnvScratch*>(smem_raw + p.off_env);
  const int lane = threadIdx.x & 31;
  const bool split = p.cl > 1;  // every CTA of the cluster evaluates the env's prologue for itself (same inputs, same result)
  const int env0 = tile_env0(p);
  const int genvs = split ? 1 : min(p.G, p.E - env0);
  const int L = p.pro_lanes, groups = 32 / L;
  const int sub = lane & (L - 1), grp = lane / L;
  int my_due = 0;
  for (int first = 0; first < genvs; first += groups) {  // warp-uniform trip count
    const int le2 = first + grp;
    const int lec = le2 < genvs ? le2 : genvs - 1;
    const bool valid = le2 < genvs && (p.env_mask == nullptr || p.env_mask[env0 + lec] != 0);
    PipeEnv unused;
    my_due |= env_prologue<false>(p, s_env[lec], unused, env0 + lec, sub, L, valid, reset, observe_only);
  }
  if (!dedicated) return my_due;
  // barrier sequence of the house warps (see step_kernel)
  if (p.solar) cta_sync();
  const int due = p.base_power_mode == MDR_BASE_INTERPOLATION ? cta_or(my_due) : (cta_sync(), 0);
  if (split) { cluster_sync_all(); cta_sync(); }
  else if (p.temp_penalty_mode != MDR_PEN_INDIVIDUAL_L2) cta_sync();
  if (due) {
    cta_sync();
    if (split) cluster_sync_all();
    cta_sync();
  }
  if (split) {
    // no thread of the cluster exits before every peer is done reading this CTA's shared memory
    if (p.metrics != nullptr && !reset) cluster_sync_all();
    cluster_sync_all();
  }
  return 0;
}

// segmented warp reduction over lanes with equal `key` (keys are contiguous runs): afterwards the
// first lane of every run holds the run's sum (fixed shuffle order => deterministic)
template <typename R>
__device__ __forceinline__ R segmented_sum(R v, int key, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const R tv = __shfl_down_sync(0xffffffffu, v, o);
    const int tk = __shfl_down_sync(0xffffffffu, key, o);
    if (lane + o < 32 && tk == key) v += tv;
  }
  return v;
}

// ----------------------------------------------------------------------------------------
// One observation row in utils.normStateDict order (utils.py:774-880) for ANY flag / neighbour mode;
// `msg_at(j)` returns the message (dT/5, sso, P/7500, Pmax/7500) of house j of the same env.
// ----------------------------------------------------------------------------------------
// Message drop of the production mode (np.random.rand() > comm_defect_prob, :992): one Philox block serves four
// messages of a house; message k takes word k % 4 of block k / 4 as a 32-bit uniform.
__device__ __forceinline__ uint4 drop_block(const KernelParams& p, unsigned h, int k4) {
  return philox4x32(h + p.house_base, (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32), STREAM_MSG + 16 * (uint32_t)k4, p.seed);
}
__device__ __forceinline__ bool drop_keep(const KernelParams& p, const uint4& r, int k) {
  const uint32_t w = (k & 3) == 0 ? r.x : (k & 3) == 1 ? r.y : (k & 3) == 2 ? r.z : r.w;
  return ((double)w + 0.5) * (1.0 / 4294967296.0) > p.comm_defect_prob;
}

template <typename R>
struct HouseRow {
  R t_air, t_mass, target, deadband, p_on, inv_lock;
  int on, lock, sso, e, li;
  unsigned h;
  double P;
};

template <typename R, typename MsgAt>
__device__ __forceinline__ void generic_row(const KernelParams& p, R* row, const HouseRow<R>& hr, const EnvScratch& es,
                                            int state_flags, int msg_flags, int comm_mode, bool has_keep, bool has_defect,
                                            int C, MsgAt msg_at) {
  const int N = p.N, half = C >> 1, li = hr.li, e = hr.e;
  const unsigned h = hr.h;
  const R inv_lock = hr.inv_lock;
  // own features, utils.normStateDict order (utils.py:774-840)
  row[0] = (hr.t_air - 20) * (R)0.2;
  row[1] = (hr.t_mass - 20) * (R)0.2;
  row[2] = (hr.target - 20) * (R)0.2;
  int c = 3;
  if (state_flags & MDR_STATE_THERMAL) row[c++] = (R)es.f_od;
  row[c++] = hr.deadband;
  if (state_flags & MDR_STATE_DAY) { row[c++] = (R)es.f_sin_day; row[c++] = (R)es.f_cos_day; }
  if (state_flags & MDR_STATE_HOUR) { row[c++] = (R)es.f_sin_hr; row[c++] = (R)es.f_cos_hr; }
  if (state_flags & MDR_STATE_SOLAR) row[c++] = (R)es.f_solar;
  row[c++] = hr.p_on * (R)p.cop_over_def_cap;
  if (state_flags & MDR_STATE_THERMAL) {
    row[c++] = (R)(p.ua[h] / p.def_ua);
    row[c++] = (R)(p.cm[h] / p.def_cm);
    row[c++] = (R)(p.ca[h] / p.def_ca);
    row[c++] = (R)(p.hm[h] / p.def_hm);
  }
  if (state_flags & MDR_STATE_HVAC) {
    row[c++] = (R)(p.hvac_cop / p.def_cop);
    row[c++] = (R)(p.hvac_latent / p.def_latent);
  }
  row[c++] = (R)hr.on;
  row[c++] = (R)hr.lock;
  row[c++] = (R)hr.sso * inv_lock;
  row[c++] = (R)1;
  row[c++] = (R)es.f_sig;
  row[c++] = (R)(hr.P * p.inv_norm_sig_agents);
  // messages, SingleHouse.message :624-662 normalised as utils.py:842-868
  R* mrow = row + c;
  const size_t tbase = comm_mode == MDR_COMM_TABLE_PER_ENV ? (size_t)e * N * C : 0;
  uint4 dr = make_uint4(0, 0, 0, 0);
  for (int k = 0; k < C; ++k) {
    int j;
    if (comm_mode == MDR_COMM_NEIGHBOURS) {
      j = k < half ? li - half + k : li + 1 + (k - half);
      if (j < 0) j += N;
      if (j >= N) j -= N;
    } else {
      j = p.comm_table[tbase + (size_t)li * C + k];
    }
    const auto m = msg_at(j);
    bool keep = true;
    if (has_keep) keep = p.msg_keep[(size_t)h * C + k] != 0;
    else if (has_defect) {
      if ((k & 3) == 0) dr = drop_block(p, h, k >> 2);
      keep = drop_keep(p, dr, k);
    }
    const R kf = keep ? (R)1 : (R)0;
    mrow[0] = m.x * kf; mrow[1] = m.y * inv_lock * kf; mrow[2] = m.z * kf; mrow[3] = m.w * kf;
    mrow += 4;
    if (msg_flags) {
      const size_t hj = (size_t)e * N + j;
      if (msg_flags & MDR_MSG_THERMAL) {
        mrow[0] = (R)(p.ua[hj] / p.def_ua) * kf; mrow[1] = (R)(p.cm[hj] / p.def_cm) * kf;
        mrow[2] = (R)(p.ca[hj] / p.def_ca) * kf; mrow[3] = (R)(p.hm[hj] / p.def_hm) * kf;
        mrow += 4;
      }
      if (msg_flags & MDR_MSG_HVAC) {
        mrow[0] = (R)(p.hvac_cop / p.def_cop) * kf; mrow[1] = (R)(p.hvac_latent / p.def_latent) * kf;
        mrow[2] = (R)(p.cap[hj] / p.def_cap) * kf;
        mrow += 3;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// the fused step kernel
// ----------------------------------------------------------------------------------------
template <typename R, int kMaxThreads, bool kFast, int kC>
__global__ void __launch_bounds__(kMaxThreads, kMaxThreads <= 256 ? MDR_BLOCKS_256 * (256 / kMaxThreads) : 1024 / kMaxThreads) step_kernel(const __grid_constant__ KernelParams p) {
  using T2 = typename Vec<R>::T2;
  using T4 = typename Vec<R>::T4;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  // kFast: implicit `neighbours` messages, default state/message flags, individual_L2 penalty, no
  // message drops -- the configuration of BASELINE configs 0-4; everything else takes the generic
  // instantiation.  kC > 0 fixes the message count at compile time (unrolled window).
  const int comm_mode = kFast ? MDR_COMM_NEIGHBOURS : p.comm_mode;
  const int state_flags = kFast ? 0 : p.state_flags;
  const int msg_flags = kFast ? 0 : p.msg_flags;
  const int pen_mode = kFast ? MDR_PEN_INDIVIDUAL_L2 : p.temp_penalty_mode;
  const bool has_keep = kFast ? false : p.msg_keep != nullptr;
  const bool has_defect = kFast ? false : p.comm_defect_prob > 0.0;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  // Programmatic dependent launch: the next step's CTAs may be scheduled while this grid runs (launch latency is what
  // bounds a 1 x 1000-house cluster); they block here until everything the previous grid wrote is visible.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // the fast instantiation is only launched for plain steps (never for reset / observe)
  const bool reset = kFast ? false : p.is_reset != 0;  // 1 = reset (grid step + obs), 2 = observe only
  const bool observe_only = kFast ? false : p.is_reset == 2;

  // ---------------- phase 0: per-env prologue, one lane per env ------------------------------
  // Normally a dedicated extra warp (no houses) runs it concurrently with the house warps'
  // loads + thermal update and then only takes part in the barriers.  When the CTA has no room
  // for an extra warp (N > 992) warp 0 runs it before loading its houses.
  if (warp >= p.house_warps) {
    prologue_warp_main(p, reset, observe_only, true);
    return;
  }
  int my_due = 0;
  if (warp == p.pro_warp) my_due = prologue_warp_main(p, reset, observe_only, false);

  const int N = p.N, C = kC > 0 ? kC : p.C;
  // Large clusters (N > 224): ONE env is split over the p.cl CTAs of a thread-block cluster; CTA `rank` owns the
  // houses [rank * cl_slice, ...) of env blockIdx.x / cl.  Cluster power / penalties cross the CTAs as per-CTA totals
  // through distributed shared memory, neighbour messages are read from the owning CTA's window.
  const int ncl = p.cl;
  const bool split = ncl > 1;
  const int rank = split ? (int)(blockIdx.x % (unsigned)ncl) : 0;
  const int slice_lo = rank * p.cl_slice;  // first house (index within the env) of this CTA
  const int env0 = tile_env0(p);
  const int genvs = split ? 1 : min(p.G, p.E - env0);
  const int H = split ? min(p.cl_slice, N - slice_lo) : genvs * N;
  const bool in_tile = tid < H;
  const int le = (in_tile && !split) ? (N == 1 ? tid : (int)__umulhi((unsigned)tid, p.div_magic)) : 0;  // tid / N
  const int li = split ? slice_lo + tid : tid - le * N;  // index of the house within its env
  const int wl = split ? tid : li;                        // ... within this CTA's message window
  const int e = env0 + le;
  // env mask (mdr_reset of a subset of the envs): houses of unselected envs are inactive
  const bool active = in_tile && (p.env_mask == nullptr || p.env_mask[e] != 0);
  const bool env_head = active && (split ? (rank == 0 && tid == 0) : li == 0);  // writes the per-env outputs
  const unsigned h = (unsigned)env0 * (unsigned)N + (unsigned)(split ? slice_lo : 0) + (unsigned)tid;
  const bool interp_mode = p.base_power_mode == MDR_BASE_INTERPOLATION;
  const bool need_pen = pen_mode != MDR_PEN_INDIVIDUAL_L2;

  T4* s_msg = reinterpret_cast<T4*>(smem_raw + p.off_msg);  // per env: [half halo | N houses | halo]
  double* s_part = reinterpret_cast<double*>(smem_raw + p.off_pw);  // [G][part_stride] warp partial power sums
  double* s_val = reinterpret_cast<double*>(smem_raw + p.off_val);
  double* s_pen = reinterpret_cast<double*>(smem_raw + p.off_pen);
  EnvScratch* s_env = reinterpret_cast<EnvScratch*>(smem_raw + p.off_env);
  R* s_stage = reinterpret_cast<R*>(smem_raw + p.off_stage);

  // ---------------- phase A loads ----------------------------------------------------------
  T2 tt = make2((R)0, (R)0);
  T4 ca4 = make4((R)0, (R)0, (R)0, (R)0), cb = ca4;
  T2 cc = make2((R)0, (R)1);
  int hv = 0, cmd = 0;
  if (active) {
    tt = reinterpret_cast<const T2*>(p.temps)[h];
    hv = p.hvac[h];
    cb = reinterpret_cast<const T4*>(p.coef_b)[h];
    cc = reinterpret_cast<const T2*>(p.coef_c)[h];
    if (!reset) {
      ca4 = reinterpret_cast<const T4*>(p.coef_a)[h];
      if (p.action_source == MDR_ACT_ARRAY) cmd = p.actions[h];
    }
  }
  if (p.solar) cta_sync();  // the thermal update needs this step's solar gain

  // ---------------- on-device GreedyMyopic controller ----------------------
  // agents/greedy_myopic_controller.py:29-49 on the state the previous step left (what its obs_dict shows):
  // houses sorted by -(T_air - target) ascending (ties: lower id first), then ONE sequential pass per env that
  // switches a house on if that keeps the total below the signal, or brings it closer and the house is not
  // locked out (the reference's operator precedence: the lockout only guards the second clause).  fp64 like pandas.
  int greedy_cmd = 0;
  if (!kFast && !reset && p.action_source == MDR_ACT_GREEDY) {
    double* g_key = s_val;                                        // [hmax] sort key, then P_on by rank
    double* g_pow = s_val + p.hmax;                               // [hmax] P_on by rank
    int* g_int = reinterpret_cast<int*>(s_val + 2 * p.hmax);      // [hmax] (house id << 1 | lockout) by rank, then commands
    const int nth = p.house_warps * 32;
    const double key = active ? -((double)tt.x - (double)cb.w) : 0.0;
    if (active) g_key[tid] = key;
    house_sync(nth);
    int rank = 0;
    if (active) {
      const double* kk = g_key + le * N;
      for (int j = 0; j < N; ++j) {
        const double kj = kk[j];
        rank += (kj < key || (kj == key && j < li)) ? 1 : 0;
      }
    }
    house_sync(nth);
    if (active) {
      g_pow[le * N + rank] = p.cap[h] / p.hvac_cop;               // obs["hvac_cooling_capacity"] / obs["hvac_COP"]
      g_int[le * N + rank] = (li << 1) | ((hv >> 1) & 1);
    }
    house_sync(nth);
    if (active && li == 0) {
      const double sig = p.signal[e];                             // obs["reg_signal"][0]
      double total = 0.0;
      int* gi = g_int + le * N;
      const double* gp = g_pow + le * N;
      for (int r = 0; r < N; ++r) {
        const double pc = gp[r];
        const int packed = gi[r];
        const bool on_r = (pc + total < sig) || (fabs(pc + total - sig) < fabs(total - sig) && !(packed & 1));
        if (on_r) total += pc;
        gi[r] = (packed & ~1) | (on_r ? 1 : 0);
      }
    }
    house_sync(nth);
    if (active) greedy_cmd = g_int[le * N + rank] & 1;
    house_sync(nth);  // s_val is reused by the interpolation refresh
  }

  // ---------------- phase A: per house -------------------------------------------------
  const R target = cb.w, p_on = cb.z, deadband = cc.x, lockdur_r = cc.y;
  int on = hv & 1, lock = (hv >> 1) & 1, sso = hv >> 2;
  R pen = 0;
  const int half = C >> 1;
  const int ns = N + C;  // shared-memory stride of one env's message window
  R pw = 0;
  if (active) {
    if (!reset) {
      if (p.action_source == MDR_ACT_ARRAY) cmd = cmd != 0;
      else if (p.action_source == MDR_ACT_BANGBANG) cmd = tt.x > target;  // agents/bangbang_controllers.py:50-61
      else if (p.action_source == MDR_ACT_GREEDY) cmd = greedy_cmd;
      else cmd = philox4x32(h + p.house_base, (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32), STREAM_ACT, p.seed).x & 1;
      // HVAC.step, :475-492
      const int dt = p.dt;
      const int lockdur = (int)lockdur_r;
      if (!on) sso += dt;
      lock = !(on || sso >= lockdur);
      const int new_on = lock ? 0 : cmd;
      if (!lock && new_on) sso = 0;
      if (!lock && !new_on && sso + dt < lockdur) lock = 1;
      on = new_on;
      // SingleHouse.update_temperature, :681-738, with the OLD outdoor temperature
      const R od_old = (R)p.od_temp[e];
      const R gain = p.solar ? (R)s_env[le].gain_now : (R)0;
      const R qa = (on ? cb.y : (R)0) + gain;
      const R tss = od_old + qa * cb.x;
      const R x = tt.x - tss, y = tt.y - tss;
      tt.x = tt.x + (ca4.x * x + ca4.y * y);
      tt.y = tt.y + (ca4.z * x + ca4.w * y);
      reinterpret_cast<T2*>(p.temps)[h] = tt;
      p.hvac[h] = (sso << 2) | (lock << 1) | on;
    }
    pw = on ? p_on : (R)0;
    const R inv_norm = (R)p.inv_norm_reg_sig;
    const T4 m = make4((tt.x - target) * (R)0.2, (R)sso, pw * inv_norm, p_on * inv_norm);
    T4* win = s_msg + le * ns;
    win[half + wl] = m;
    if (!split) {
      if (li < C - half) win[half + N + li] = m;        // wrap-around halo after the last house
      if (li >= N - half) win[li - (N - half)] = m;     // ... and before the first one
    }
    // utils.deadbandL2, utils.py:1266-1274
    const R hi = target + deadband / 2, lo = target - deadband / 2;
    if (hi < tt.x) pen = (tt.x - hi) * (tt.x - hi);
    else if (lo > tt.x) pen = (lo - tt.x) * (lo - tt.x);
    if (need_pen) s_pen[tid] = (double)pen;
  }
  // cluster power: per-warp segmented partial sums (keyed by env), summed per thread after the barrier
  {
    const int key = active ? le : -1;
    const double part = (double)segmented_sum<R>(pw, key, lane);
    const int prev_key = __shfl_up_sync(0xffffffffu, key, 1);
    if (active && (lane == 0 || prev_key != key)) {
      const int first_warp = (le * N) >> 5;
      s_part[le * p.part_stride + (warp - first_warp)] = part;
    }
  }
  // Fast path with a dedicated prologue warp: the house warps first meet on their own barrier (1),
  // sum the power and assemble their observation rows, and only then join the prologue warp on
  // barrier 0 -- its fp64 latency is hidden behind the row assembly.
  const bool deferred = kFast && !split && p.pro_warp >= p.house_warps && p.obs != nullptr;
  int any_due = 0;
  if (deferred) house_sync(p.house_warps * 32);
  else any_due = interp_mode ? cta_or(my_due) : (cta_sync(), 0);

  // every thread now knows its env's power: sum the warp partials in warp order
  double P = 0.0;
  if (split) {
    // per-CTA totals (power; sum and max of the penalties) -> cluster barrier -> warp 0 adds the peers' totals in
    // rank order (deterministic) -> one more CTA barrier publishes them to the house threads
    ClusterTot* s_tot = reinterpret_cast<ClusterTot*>(smem_raw + p.off_cl);
    if (warp == 0) {
      const int nw = (H + 31) >> 5;
      double ps = 0.0, pn = 0.0, pm = 0.0;
      for (int w = lane; w < nw; w += 32) ps += s_part[w];
      if (need_pen)
        for (int i = lane; i < H; i += 32) {
          const double v = s_pen[i];
          pn += v / N;
          pm = fmax(pm, v);
        }
      ps = warp_sum(ps); pn = warp_sum(pn); pm = warp_max(pm);
      if (lane == 0) { s_tot->P = ps; s_tot->pen_sum = pn; s_tot->pen_max = pm; }
    }
    cluster_sync_all();
    if (warp == 0) {
      double tp = 0.0, tn = 0.0, tm = 0.0;
      if (lane < ncl) {
        const ClusterTot* peer = dsmem_peer(s_tot, lane);
        tp = peer->P; tn = peer->pen_sum; tm = peer->pen_max;
      }
      double sp = 0.0, sn = 0.0, sm = 0.0;
      for (int r = 0; r < ncl; ++r) {
        sp += __shfl_sync(0xffffffffu, tp, r);
        sn += __shfl_sync(0xffffffffu, tn, r);
        sm = fmax(sm, __shfl_sync(0xffffffffu, tm, r));
      }
      if (lane == 0) { s_env[0].P = sp; s_env[0].pen_mean = sn; s_env[0].pen_max = sm; }
    }
    cta_sync();
    P = s_env[0].P;
  } else if (active) {
    const int first_warp = (le * N) >> 5, last_warp = (le * N + N - 1) >> 5;
    for (int w = 0; w <= last_warp - first_warp; ++w) P += s_part[le * p.part_stride + w];
  }
  // message of house j (index within the env) of this thread's env: local window, or the owning CTA's (DSMEM)
  auto msg_at = [&](int j) -> T4 {
    if (!split) return s_msg[le * ns + half + j];
    const int rj = j / p.cl_slice;
    const T4* src = s_msg + half + (j - rj * p.cl_slice);
    if (rj != rank) src = dsmem_peer(src, rj);
    return *src;
  };

  const int F = p.F;
  const int rpp = p.rows_per_pass;
  R* stage = s_stage + warp * rpp * F;
  const int wrow0 = warp * 32;
  const int nrows_w = max(0, min(32, H - wrow0));
  const R inv_lock = inv_real(lockdur_r);
  // fast-path row: [T_air, T_mass, target, deadband, cap, on, lockout, sso, 1, signal, power | C x 4 messages]
  auto fast_row = [&](R* row) {
    row[0] = (tt.x - 20) * (R)0.2;
    row[1] = (tt.y - 20) * (R)0.2;
    row[2] = (target - 20) * (R)0.2;
    row[3] = deadband;
    row[4] = p_on * (R)p.cop_over_def_cap;
    row[5] = (R)on;
    row[6] = (R)lock;
    row[7] = (R)sso * inv_lock;
    row[8] = (R)1;
    row[10] = (R)(P * p.inv_norm_sig_agents);
    // neighbours (:816-828) = the C window entries around this house, skipping itself
    const T4* win = s_msg + le * ns + li;
    R* mrow = row + 11;
#pragma unroll
    for (int k = 0; k < (kC > 0 ? kC : C); ++k) {
      T4 m;
      if (!split) m = win[k + (k >= half ? 1 : 0)];
      else {
        int j = li - half + k + (k >= half ? 1 : 0);
        if (j < 0) j += N;
        if (j >= N) j -= N;
        m = msg_at(j);
      }
      mrow[4 * k + 0] = m.x;
      mrow[4 * k + 1] = m.y * inv_lock;
      mrow[4 * k + 2] = m.z;
      mrow[4 * k + 3] = m.w;
    }
  };
  if (deferred) {
    if (lane < min(rpp, nrows_w)) fast_row(stage + lane * F);
    any_due = interp_mode ? cta_or(my_due) : (cta_sync(), 0);
  }

  // ---------------- phase B (generic penalty modes only): mean / max of the penalties --------
  if (need_pen && !split) {
    const int nwarps = p.house_warps;
    for (int le2 = warp; le2 < genvs; le2 += nwarps) {
      double pmean = 0.0, pmax = 0.0;
      for (int i = lane; i < N; i += 32) {
        const double v = s_pen[le2 * N + i];
        pmean += v / N;
        pmax = fmax(pmax, v);
      }
      pmean = warp_sum(pmean);
      pmax = warp_max(pmax);
      if (lane == 0) {
        s_env[le2].pen_mean = pmean;
        s_env[le2].pen_max = pmax;
      }
    }
    cta_sync();
  }

  // per-env state written back by the env's first house thread
  if (env_head && !observe_only) {
    const EnvScratch& es = s_env[le];
    p.cluster_power[e] = P;
    if (!reset) {
      p.od_temp[e] = es.od_new;
      p.t_epoch[e] = (int64_t)es.t_new;
    }
    if (p.solar) p.solar_gain[e] = es.gain_now;
    if (!es.due) {
      p.base_power[e] = es.base;
      p.signal[e] = es.sig_new;
      if (interp_mode) p.time_since_interp[e] = es.tsi;
    }
  }

  // ---------------- phases C/D: interpolation refresh (every interp_update_period) -------
  if (any_due) {
    const int nb = p.interp_nb_agents;
    const int nsamp = N <= nb ? N : nb;
    // split env: the CTA of rank 0 evaluates all samples (the peers' new temperatures are visible in global memory
    // since the cluster barrier) and hands the new signal to the peers
    if (active && s_env[le].due && li < nsamp && rank == 0) {
      int src = li;
      if (N > nb) {
        if (p.interp_ids) src = p.interp_ids[(size_t)e * nb + li];
        else {
          const uint4 r = philox4x32((uint32_t)(e + p.env_base), (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32),
                                     STREAM_IDS + 16 * (uint32_t)li, p.seed);
          src = (int)(((uint64_t)r.x * (uint64_t)N) >> 32);
        }
      }
      const size_t hs = (size_t)e * N + src;
      const T2 t2 = reinterpret_cast<const T2*>(p.temps)[hs];
      const double tg = (double)reinterpret_cast<const T4*>(p.coef_b)[hs].w;
      const EnvScratch& es = s_env[le];
      s_val[tid] = interp_eval<R>(p, p.interp_key[hs], (double)t2.x - tg, (double)t2.y - tg, es.od_new - tg, es.hour_s,
                                  es.date);
    }
    cta_sync();
    if (env_head && s_env[le].due) {
      EnvScratch& es = s_env[le];
      double base = 0.0;
      for (int i = 0; i < nsamp; ++i) base = add_rn(base, s_val[le * N + i]);  // id order, :1218-1232
      if (N > nb) base = mul_rn(base, (double)N / (double)nb);
      const double sig = grid_signal(p, base, es.time_sec, es.sig_noise, p.artificial_ratio[e], p.max_power[e]);
      p.base_power[e] = base;
      p.time_since_interp[e] = 0;
      p.signal[e] = sig;
      es.f_sig = sig * p.inv_norm_sig_agents;
      es.sig_new = sig;
    }
    if (split) {
      cluster_sync_all();
      if (rank != 0 && tid == 0 && s_env[0].due) {
        const EnvScratch* lead = dsmem_peer(s_env, 0);
        s_env[0].f_sig = lead->f_sig;
        s_env[0].sig_new = lead->sig_new;
      }
    }
    cta_sync();
  }

  // ---------------- phase E: reward + observation ----------------------------------------
  R reward_r = 0;
  if (active && !reset && (p.reward != nullptr || p.metrics != nullptr)) {
    const EnvScratch& es = s_env[le];
    double tp = (double)pen;
    if (pen_mode == MDR_PEN_COMMON_L2) tp = es.pen_mean;
    else if (pen_mode == MDR_PEN_COMMON_MAX) tp = es.pen_max;
    else if (pen_mode == MDR_PEN_MIXTURE)
      tp = (p.mix_alpha_ind * tp + p.mix_alpha_common * es.pen_mean + p.mix_alpha_max * es.pen_max) /
           (p.mix_alpha_ind + p.mix_alpha_common + p.mix_alpha_max);
    // reg_signal_penalty :244-247 with the OLD signal; weighting :364-372
    const double dn = (P - es.s_old) * p.inv_n;
    reward_r = (R)(-(tp * p.k_temp + dn * dn * p.k_sig));
    if (p.reward != nullptr) reinterpret_cast<R*>(p.reward)[h] = reward_r;
  }

  // ---------------- metric accumulators for single-step launches of ANY configuration ----------
  // (array actions, interpolation, every penalty mode): the quantities main-deploy.py:124-149 and
  // metrics.py:22-30 accumulate, reduced per env in a fixed order and += into MdrEnvs.metrics
  if (!kFast && !reset && p.metrics != nullptr) {
    double* s_met = reinterpret_cast<double*>(smem_raw + p.off_met);  // [G][part_stride][5]
    const int key = active ? le : -1;
    const int prev_key = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = active && (lane == 0 || prev_key != key);
    const double err = active ? (double)tt.x - (double)target : 0.0;
    const double v0 = segmented_sum<double>(active ? (double)reward_r : 0.0, key, lane);
    const double v1 = segmented_sum<double>(err, key, lane), v2 = segmented_sum<double>(fabs(err), key, lane);
    const double v3 = segmented_sum<double>(err * err, key, lane);
    double v4 = fabs(err);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double tv = __shfl_down_sync(0xffffffffu, v4, o);
      const int tk = __shfl_down_sync(0xffffffffu, key, o);
      if (lane + o < 32 && tk == key) v4 = fmax(v4, tv);
    }
    const int first_warp = (le * N) >> 5;
    if (head) {
      double* d = s_met + (size_t)(le * p.part_stride + (warp - first_warp)) * 5;
      d[0] = v0; d[1] = v1; d[2] = v2; d[3] = v3; d[4] = v4;
    }
    house_sync(p.house_warps * 32);
    if (split) {
      // per-CTA totals in warp order -> cluster barrier -> the env's first thread adds them in rank order
      ClusterTot* s_tot = reinterpret_cast<ClusterTot*>(smem_raw + p.off_cl);
      if (tid == 0) {
        const int nw = (H + 31) >> 5;
        double t[4] = {0.0, 0.0, 0.0, 0.0}, mx = 0.0;
        for (int w = 0; w < nw; ++w) {
          const double* d = s_met + (size_t)w * 5;
          for (int k = 0; k < 4; ++k) t[k] += d[k];
          mx = fmax(mx, d[4]);
        }
        for (int k = 0; k < 4; ++k) s_tot->met[k] = t[k];
        s_tot->met[4] = mx;
      }
      cluster_sync_all();
    }
    if (env_head) {
      double t[4] = {0.0, 0.0, 0.0, 0.0}, mx = 0.0;
      if (split) {
        const ClusterTot* s_tot = reinterpret_cast<const ClusterTot*>(smem_raw + p.off_cl);
        for (int r = 0; r < ncl; ++r) {
          const ClusterTot* peer = dsmem_peer(s_tot, r);
          for (int k = 0; k < 4; ++k) t[k] += peer->met[k];
          mx = fmax(mx, peer->met[4]);
        }
      } else {
        const int nparts = ((le * N + N - 1) >> 5) - first_warp + 1;
        for (int w = 0; w < nparts; ++w) {
          const double* d = s_met + (size_t)(le * p.part_stride + w) * 5;
          for (int k = 0; k < 4; ++k) t[k] += d[k];
          mx = fmax(mx, d[4]);
        }
      }
      const EnvScratch& es = s_env[le];
      const double dsp = es.sig_new - P;  // NEW signal (after a refresh, if one was due) minus this step's consumption
      double* m = p.metrics + (size_t)e * MDR_N_METRICS;
      m[MDR_M_STEPS] += 1.0;
      m[MDR_M_SUM_MEAN_REWARD] += t[0] * p.inv_n;
      m[MDR_M_SUM_MEAN_TEMP_OFFSET] += t[1] * p.inv_n;
      m[MDR_M_SUM_MEAN_TEMP_ERROR] += t[2] * p.inv_n;
      m[MDR_M_SUM_SQ_TEMP_ERROR] += t[3];
      m[MDR_M_SUM_SQ_MAX_TEMP_ERROR] += mx * mx;
      m[MDR_M_MAX_TEMP_ERROR] = fmax(m[MDR_M_MAX_TEMP_ERROR], mx);
      m[MDR_M_SUM_OD_TEMP] += es.od_new;
      m[MDR_M_SUM_SIGNAL] += es.sig_new;
      m[MDR_M_SUM_CONSUMPTION] += P;
      m[MDR_M_SUM_SIGNAL_OFFSET] += dsp;
      m[MDR_M_SUM_SIGNAL_ERROR] += fabs(dsp);
      m[MDR_M_SUM_SQ_SIGNAL_ERROR] += dsp * dsp;
    }
  }

  if (p.obs != nullptr) {
  R* gobs = reinterpret_cast<R*>(p.obs) + (size_t)((unsigned)env0 * (unsigned)N + (unsigned)(split ? slice_lo : 0) + (unsigned)wrow0) * F;
  bool issued = false;
  for (int pass0 = 0; pass0 < nrows_w; pass0 += rpp) {
    const int nr = min(rpp, nrows_w - pass0);
    if (kFast) {
      if (lane >= pass0 && lane < pass0 + nr) {
        R* row = stage + (lane - pass0) * F;
        if (!(deferred && pass0 == 0)) fast_row(row);
        row[9] = (R)s_env[le].f_sig;
      }
    } else if (lane >= pass0 && lane < pass0 + nr) {
      HouseRow<R> hr;
      hr.t_air = tt.x; hr.t_mass = tt.y; hr.target = target; hr.deadband = deadband; hr.p_on = p_on; hr.inv_lock = inv_lock;
      hr.on = on; hr.lock = lock; hr.sso = sso; hr.P = P; hr.h = h; hr.e = e; hr.li = li;
      generic_row<R>(p, stage + (lane - pass0) * F, hr, s_env[le], state_flags, msg_flags, comm_mode, has_keep, has_defect, C, msg_at);
    }
    R* dst = gobs + (size_t)pass0 * F;
    const uint32_t bytes = (uint32_t)(nr * F * sizeof(R));
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(dst) | bytes) & 15) == 0;
    if (bulk_ok) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) bulk_store_s2g(dst, stage, bytes);
      issued = true;
      if (pass0 + rpp < nrows_w) {
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
      }
    } else {
      __syncwarp();
      for (int i = lane; i < nr * F; i += 32) dst[i] = stage[i];
      __syncwarp();
    }
  }
  if (issued && lane == 0) bulk_wait_read_all();
  }
  // no CTA of the cluster may exit while a peer still reads its shared memory
  if (split) cluster_sync_all();
}

#include "mdr_pipe.cuh"
#include "mdr_pipe_split.cuh"
#include "mdr_fused.cuh"
#include "mdr_populate.cuh"
#include "mdr_big.cuh"
#include "mdr_rollout.cuh"
#include "mdr_compact.cuh"

// ----------------------------------------------------------------------------------------
// host-side launch helpers
// ----------------------------------------------------------------------------------------
template <typename R>
static cudaError_t launch_precompute(const KernelParams& kp, cudaStream_t stream) {
  const size_t total = (size_t)kp.E * kp.N;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  precompute_kernel<R><<<blocks, threads, 0, stream>>>(kp);
  return cudaGetLastError();
}

// Opt-in to the full shared-memory carve-out, once per (kernel, device).  The latch is an atomic bit per device:
// two host threads racing on the first launch both set the (idempotent) attribute.
template <typename K>
static cudaError_t ensure_max_smem(K kernel, std::atomic<uint64_t>& latch) {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  const uint64_t bit = 1ull << (dev & 63);
  if (dev < 64 && (latch.load(std::memory_order_acquire) & bit)) return cudaSuccess;
  err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MDR_MAX_SMEM_BYTES);
  if (err != cudaSuccess) return err;
  if (dev < 64) latch.fetch_or(bit, std::memory_order_release);
  return cudaSuccess;
}

template <typename R, int kMaxThreads, bool kFast, int kC>
static cudaError_t launch_step_t(const KernelParams& kp, const Geometry& g, cudaStream_t stream) {
  static std::atomic<uint64_t> latch{0};
  cudaError_t err = ensure_max_smem(step_kernel<R, kMaxThreads, kFast, kC>, latch);
  if (err != cudaSuccess) return err;
  if (g.cluster > 8) {  // clusters larger than 8 CTAs are "non-portable": opt in once per device
    static std::atomic<uint64_t> np_latch{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (!(np_latch.load(std::memory_order_acquire) & bit)) {
      err = cudaFuncSetAttribute(step_kernel<R, kMaxThreads, kFast, kC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (err != cudaSuccess) return err;
      np_latch.fetch_or(bit, std::memory_order_release);
    }
  }
  cudaLaunchAttribute attrs[2];
  int n_attrs = 0;
  if (g.cluster > 1) {  // one env per thread-block cluster
    attrs[n_attrs].id = cudaLaunchAttributeClusterDimension;
    attrs[n_attrs].val.clusterDim.x = (unsigned)g.cluster;
    attrs[n_attrs].val.clusterDim.y = 1;
    attrs[n_attrs].val.clusterDim.z = 1;
    ++n_attrs;
  }
  if (!g.no_pdl) {
    attrs[n_attrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n_attrs].val.programmaticStreamSerializationAllowed = 1;
    ++n_attrs;
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)g.ctas);
  lc.blockDim = dim3((unsigned)g.threads);
  lc.dynamicSmemBytes = g.smem_bytes;
  lc.stream = stream;
  lc.attrs = attrs;
  lc.numAttrs = n_attrs;
  return cudaLaunchKernelEx(&lc, step_kernel<R, kMaxThreads, kFast, kC>, kp);
}

template <typename R, bool kFast, int kC>
static cudaError_t launch_step_f(const KernelParams& kp, const Geometry& g, cudaStream_t stream) {
  if (g.threads <= 128) return launch_step_t<R, 128, kFast, kC>(kp, g, stream);
  if (g.threads <= 256) return launch_step_t<R, 256, kFast, kC>(kp, g, stream);
  if (g.threads <= 512) return launch_step_t<R, 512, kFast, kC>(kp, g, stream);
  return launch_step_t<R, 1024, kFast, kC>(kp, g, stream);
}

template <typename R>
static cudaError_t launch_step_r(const KernelParams& kp, const Geometry& g, cudaStream_t stream) {
  const bool fast = kp.is_reset == 0 && kp.comm_mode == MDR_COMM_NEIGHBOURS && kp.state_flags == 0 && kp.msg_flags == 0 &&
                    kp.temp_penalty_mode == MDR_PEN_INDIVIDUAL_L2 && kp.msg_keep == nullptr &&
                    !(kp.comm_defect_prob > 0.0) && kp.action_source != MDR_ACT_GREEDY && kp.metrics == nullptr;
  if (fast && kp.C == 10) return launch_step_f<R, true, 10>(kp, g, stream);
  return fast ? launch_step_f<R, true, 0>(kp, g, stream) : launch_step_f<R, false, 0>(kp, g, stream);
}

// Resident CTAs per SM of one pipelined instantiation, cached per (device, threads, shared memory) under a mutex
// (different geometries alternate freely; any host thread may launch).
struct OccEntry { int dev, threads; size_t smem; int ctas_per_sm, sm_count; };

template <int kC, int kAct, bool kObs, int kVar, bool kDyn>
static cudaError_t launch_pipe_t(const KernelParams& kp_in, const Geometry& g, cudaStream_t stream) {
  static std::atomic<uint64_t> latch{0};
  static std::mutex mu;
  static std::vector<OccEntry> cache;
  cudaError_t err = ensure_max_smem(step_pipe_kernel<kC, kAct, kObs, kVar, kDyn>, latch);
  if (err != cudaSuccess) return err;
  int dev = 0;
  err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  int ctas_per_sm = 0, sm_count = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    for (const OccEntry& o : cache)
      if (o.dev == dev && o.threads == g.threads && o.smem == g.pipe_smem_bytes) { ctas_per_sm = o.ctas_per_sm; sm_count = o.sm_count; }
    if (ctas_per_sm == 0) {
      err = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      if (err != cudaSuccess) return err;
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, step_pipe_kernel<kC, kAct, kObs, kVar, kDyn>, g.threads, g.pipe_smem_bytes);
      if (err != cudaSuccess) return err;
      if (ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
      cache.push_back(OccEntry{dev, g.threads, g.pipe_smem_bytes, ctas_per_sm, sm_count});
    }
  }
  KernelParams kp = kp_in;
  kp.n_tiles = g.ctas;
  int grid = sm_count * ctas_per_sm;
  if (g.max_ctas > 0 && grid > g.max_ctas) grid = g.max_ctas;
  if (grid > g.ctas) grid = g.ctas;
  // Programmatic dependent launch: the CTAs of step t+1 may become resident (and set up their mbarriers and index
  // arithmetic) while the last CTAs of step t are still draining; they block in griddepcontrol.wait before touching
  // global memory, so the data dependency on the whole previous grid is unchanged.
  cudaLaunchAttribute attrs[2];
  int n_attrs = 0;
  if (!g.no_pdl) {
    attrs[n_attrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n_attrs].val.programmaticStreamSerializationAllowed = 1;
    ++n_attrs;
  }
  if (g.l2_window_bytes != 0) {
    // keep the per-house state/coefficients L2-resident across steps (opt-in)
    cudaLaunchAttribute& attr = attrs[n_attrs++];
    attr.id = cudaLaunchAttributeAccessPolicyWindow;
    attr.val.accessPolicyWindow.base_ptr = const_cast<void*>(g.l2_window_base);
    attr.val.accessPolicyWindow.num_bytes = g.l2_window_bytes;
    attr.val.accessPolicyWindow.hitRatio = g.l2_hit_ratio;
    attr.val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  }
  // In-order tile claiming pays a grid barrier behind the per-env prologue for a balanced, shorter tile loop: worth
  // it from about a dozen tiles per CTA (16 384 x 100: 18), not for small problems (4 096 x 50: 4.6 tiles per CTA).
  // MDR_DYN_MIN_TILES overrides the threshold (tiles per CTA).
  static const int dyn_min = [] { const char* e = getenv("MDR_DYN_MIN_TILES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 14; }();
  if (kDyn && (long long)kp.n_tiles < (long long)dyn_min * grid) {
    kp.dyn_off = 0;
    return launch_pipe_t<kC, kAct, kObs, kVar, false>(kp, g, stream);
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(g.threads);
  lc.dynamicSmemBytes = g.pipe_smem_bytes;
  lc.stream = stream;
  lc.attrs = attrs;
  lc.numAttrs = n_attrs;
  return cudaLaunchKernelEx(&lc, step_pipe_kernel<kC, kAct, kObs, kVar, kDyn>, kp);
}

template <int kC, bool kObs, int kVar>
static cudaError_t launch_pipe_c(const KernelParams& kp, const Geometry& g, cudaStream_t stream) {
  // (in-order tile claiming is instantiated for kernels that write observations: without them the tiles are short and
  //  the per-env prologue, not the tile loop, bounds the launch)
  if (kObs && kp.dyn_off != 0) {
    if (kp.action_source == MDR_ACT_ARRAY) return launch_pipe_t<kC, MDR_ACT_ARRAY, kObs, kVar, kObs>(kp, g, stream);
    if (kp.action_source == MDR_ACT_BANGBANG) return launch_pipe_t<kC, MDR_ACT_BANGBANG, kObs, kVar, kObs>(kp, g, stream);
    return launch_pipe_t<kC, MDR_ACT_RANDOM, kObs, kVar, kObs>(kp, g, stream);
  }
  KernelParams ks = kp;
  ks.dyn_off = 0;
  if (kp.action_source == MDR_ACT_ARRAY) return launch_pipe_t<kC, MDR_ACT_ARRAY, kObs, kVar, false>(ks, g, stream);
  if (kp.action_source == MDR_ACT_BANGBANG) return launch_pipe_t<kC, MDR_ACT_BANGBANG, kObs, kVar, false>(ks, g, stream);
  return launch_pipe_t<kC, MDR_ACT_RANDOM, kObs, kVar, false>(ks, g, stream);
}

template <int kC, bool kObs>
static cudaError_t launch_pipe_v(const KernelParams& kp, const Geometry& g, cudaStream_t stream) {
  // variant bits: 1 = metric accumulators, 2 = message drops (only meaningful with an observation)
  const bool drops = kObs && (kp.msg_keep != nullptr || kp.comm_defect_prob > 0.0);
  const bool metrics = kp.metrics != nullptr;
  if (kObs && drops) return metrics ? launch_pipe_c<kC, kObs, (kObs ? 3 : 1)>(kp, g, stream) : launch_pipe_c<kC, kObs, (kObs ? 2 : 0)>(kp, g, stream);
  return metrics ? launch_pipe_c<kC, kObs, 1>(kp, g, stream) : launch_pipe_c<kC, kObs, 0>(kp, g, stream);
}

cudaError_t launch_pipe(const KernelParams& kp, const Geometry& g, cudaStream_t stream) {
  if (kp.obs != nullptr) return kp.C == 10 ? launch_pipe_v<10, true>(kp, g, stream) : launch_pipe_v<0, true>(kp, g, stream);
  return kp.C == 10 ? launch_pipe_v<10, false>(kp, g, stream) : launch_pipe_v<0, false>(kp, g, stream);
}

#include "mdr_pipe_split_host.cuh"
#include "mdr_wide.cuh"

// tiles per prologue pass (power of two): as deep as the lanes (one per env at least), the ring
// (kMaxRing slots) and the shared-memory budget (the observation staging tile is the big consumer)
// allow.  MDR_PRO_BATCH overrides the cap for tuning.
int pipe_pro_batch(int envs_per_cta, bool has_obs) {
  int cap = has_obs ? 4 : 8;
  static const int tuned = [] {  // read once (thread-safe static initialisation)
    const char* s = getenv("MDR_PRO_BATCH");
    const int v = s ? atoi(s) : 0;
    return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 0;
  }();
  if (tuned) cap = tuned;
  int b = 1;
  while (2 * b <= cap && 2 * b * envs_per_cta <= 32) b *= 2;
  return b;
}

// The pipelined kernel takes the default observation layout (implicit `neighbours` messages, no optional feature
// blocks), the individual_L2 penalty and whole envs of <= 224 houses per tile; solar gain, message drops (replayed or
// Philox) and the metric accumulators are variants of it (runtime / template), so none of them changes the kernel.
bool pipe_eligible(const KernelParams& kp, const Geometry& g, int precision) {
  return precision == MDR_F32 && kp.is_reset == 0 && kp.comm_mode == MDR_COMM_NEIGHBOURS && kp.state_flags == 0 &&
         kp.msg_flags == 0 && kp.temp_penalty_mode == MDR_PEN_INDIVIDUAL_L2 && g.pro_warp >= g.house_warps &&
         g.threads <= 256 && g.rows_per_pass == 32 && g.pipe_smem_bytes > 0 && g.cluster <= 1 &&
         kp.action_source != MDR_ACT_GREEDY;
}

cudaError_t launch_precompute_any(const KernelParams& kp, int precision, cudaStream_t stream) {
  return precision == MDR_F32 ? launch_precompute<float>(kp, stream) : launch_precompute<double>(kp, stream);
}

cudaError_t launch_step_any(const KernelParams& kp, const Geometry& g, int precision, cudaStream_t stream) {
  return precision == MDR_F32 ? launch_step_r<float>(kp, g, stream) : launch_step_r<double>(kp, g, stream);
}

// layout of the pipelined kernel: message window and power partials double buffered, a ring of PipeEnv records, one
// contiguous staging tile for the tile's G*N observation rows, two cp.async input stages
size_t pipe_smem_layout(KernelParams* kp, int hmax, int genvs, int n_houses, int n_features, bool need_val, bool has_obs,
                        int n_comm, int part_stride, int pro_batch) {
  size_t o = 0;
  const size_t off_msg = o;   o += align16((size_t)2 * genvs * (n_houses + n_comm) * 4 * sizeof(float));
  const size_t off_pw = o;    o += align16((size_t)2 * genvs * part_stride * sizeof(double));
  const size_t off_met = o;   o += align16((size_t)2 * genvs * part_stride * 5 * sizeof(float));  // metric partials
  const size_t off_val = o;   o += need_val ? align16(((size_t)hmax + (size_t)genvs * 40) * sizeof(double)) : 0;  // values | per-env partial sums | per-env scalars
  const size_t off_grid = o;  o += need_val ? align16(sizeof(InterpGrid)) : 0;
  const size_t off_env = o;   o += align16((size_t)(pro_batch < 2 ? 4 : 2 * pro_batch) * genvs * sizeof(PipeEnv));  // (in-order claiming: 4 slots)
  const size_t off_ctl = o;   o += align16(sizeof(PipeCtl));
  const size_t off_stage = o; o += has_obs ? align16((size_t)genvs * n_houses * n_features * sizeof(float)) : 0;
  const size_t off_in = o;    o += align16((size_t)2 * hmax * 52);
  if (kp) {
    kp->off_msg = (int)off_msg; kp->off_pw = (int)off_pw; kp->off_val = (int)off_val; kp->off_pen = 0;
    kp->off_env = (int)off_env; kp->off_stage = (int)off_stage; kp->off_in = (int)off_in; kp->off_ctl = (int)off_ctl;
    kp->off_met = (int)off_met; kp->off_grid = (int)off_grid;
  }
  return o;
}

size_t step_smem_layout(KernelParams* kp, int real_bytes, int hmax, int genvs, int nwarps, int rows_per_pass,
                        int n_features, bool need_val, bool need_pen, bool has_obs, int n_comm, int part_stride, bool need_met,
                        int part_slots) {
  const SmemLayout L = smem_layout(real_bytes, hmax, genvs, nwarps, rows_per_pass, n_features, need_val, need_pen, has_obs,
                                   n_comm, part_stride, need_met, part_slots);
  if (kp) {
    kp->off_msg = (int)L.off_msg; kp->off_pw = (int)L.off_pw; kp->off_val = (int)L.off_val;
    kp->off_pen = (int)L.off_pen; kp->off_env = (int)L.off_env; kp->off_stage = (int)L.off_stage;
    kp->off_met = (int)L.off_met; kp->off_cl = (int)L.off_cl;
  }
  return L.total;
}

}  // namespace mdr

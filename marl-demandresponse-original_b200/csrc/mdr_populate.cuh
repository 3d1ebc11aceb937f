// Device-side population draw (reset-time randomness) and its launcher.
// Included by mdr_kernels.cu inside namespace mdr (after the shared device helpers); not a standalone translation unit.
#pragma once

// ----------------------------------------------------------------------------------------
// Device-side population draw: the reset-time randomness of
// utils.applyPropertyNoise (utils.py:573-709), HVAC.__init__ (:430-434), ClusterHouses.__init__
// (:789-793) and PowerGrid.__init__ (:1116, :1182-1184) from counter-based Philox streams keyed by
// (house or env, draw_index) -- distribution-level (not bit-level) parity with python's `random`.
// One CTA per env; an optional env mask re-draws only some envs (partial reset) and leaves every
// byte of the others untouched.
// ----------------------------------------------------------------------------------------
enum : uint32_t { STREAM_POP_HOUSE = 6, STREAM_POP_ENV = 7 };

__device__ __forceinline__ double gauss01(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {  // Box-Muller, fp64
  const double u1 = u01(a, b), u2 = u01(c, d);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
// random.triangular(low, high, mode) as CPython implements it
__device__ __forceinline__ double triangular(double u, double low, double high, double mode) {
  if (high == low) return low;
  double c = (mode - low) / (high - low);
  if (u > c) {
    u = 1.0 - u;
    c = 1.0 - c;
    const double t = low; low = high; high = t;
  }
  return low + (high - low) * sqrt(u * c);
}

__global__ void __launch_bounds__(128) populate_kernel(const __grid_constant__ KernelParams p, const MdrPopulationSpec s,
                                                       const uint8_t* __restrict__ env_mask, double* raw_ua, double* raw_cm,
                                                       double* raw_ca, double* raw_hm, double* raw_cap, double* raw_target,
                                                       double* raw_deadband, int32_t* lockout_dur, uint64_t draw_index) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_cap = reinterpret_cast<double*>(smem_raw);  // [N] for the id-ordered max_power sum (:796-802)
  const int e = blockIdx.x;
  if (env_mask != nullptr && env_mask[e] == 0) return;
  const int N = p.N;
  const uint32_t d_lo = (uint32_t)draw_index, d_hi = (uint32_t)(draw_index >> 32);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const unsigned h = (unsigned)e * (unsigned)N + (unsigned)i;
    const uint4 r0 = philox4x32(h, d_lo, d_hi, STREAM_POP_HOUSE, p.seed);
    const uint4 r1 = philox4x32(h, d_lo, d_hi, STREAM_POP_HOUSE + 16, p.seed);
    const uint4 r2 = philox4x32(h, d_lo, d_hi, STREAM_POP_HOUSE + 32, p.seed);
    const uint4 r3 = philox4x32(h, d_lo, d_hi, STREAM_POP_HOUSE + 48, p.seed);
    const uint4 r4 = philox4x32(h, d_lo, d_hi, STREAM_POP_HOUSE + 64, p.seed);
    // apply_house_noise, utils.py:623-666
    const double t_air = s.init_air_temp + fabs(s.std_start_temp * gauss01(r0.x, r0.y, r0.z, r0.w));
    const double t_mass = s.init_mass_temp + fabs(s.std_start_temp * gauss01(r1.x, r1.y, r1.z, r1.w));
    const double target = s.target_temp + fabs(s.std_target_temp * gauss01(r2.x, r2.y, r2.z, r2.w));
    const double lo = s.factor_thermo_low, hi = s.factor_thermo_high;
    raw_ua[h] = s.ua * triangular(u01(r3.x, r3.y), lo, hi, 1.0);
    raw_cm[h] = s.cm * triangular(u01(r3.z, r3.w), lo, hi, 1.0);
    raw_ca[h] = s.ca * triangular(u01(r4.x, r4.y), lo, hi, 1.0);
    const uint4 r5 = philox4x32(h, d_lo, d_hi, STREAM_POP_HOUSE + 80, p.seed);
    raw_hm[h] = s.hm * triangular(u01(r4.z, r4.w), lo, hi, 1.0);
    raw_target[h] = target;
    raw_deadband[h] = s.deadband;
    // apply_hvac_noise (random.choices of the capacity list), utils.py:669-676
    const int ncap = s.n_cap > 0 ? s.n_cap : 1;
    const double cap = s.cap_list[min(ncap - 1, (int)(u01(r5.x, r5.y) * ncap))];
    raw_cap[h] = cap;
    s_cap[i] = cap;
    // HVAC.__init__ lockout noise: randint(-noise, +noise), :430-434
    const int span = 2 * s.lockout_noise + 1;
    const int dur = s.lockout_duration - s.lockout_noise + min(span - 1, (int)(u01(r5.z, r5.w) * span));
    lockout_dur[h] = dur;
    if (p.temps != nullptr) {
      if (p.precision == MDR_F32) reinterpret_cast<float2*>(p.temps)[h] = make_float2((float)t_air, (float)t_mass);
      else reinterpret_cast<double2*>(p.temps)[h] = make_double2(t_air, t_mass);
    }
    p.hvac[h] = dur << 2;  // off, not locked out, seconds_since_off = lockout duration (:433)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint4 q0 = philox4x32((uint32_t)e, d_lo, d_hi, STREAM_POP_ENV, p.seed);
    const uint4 q1 = philox4x32((uint32_t)e, d_lo, d_hi, STREAM_POP_ENV + 16, p.seed);
    const uint4 q2 = philox4x32((uint32_t)e, d_lo, d_hi, STREAM_POP_ENV + 32, p.seed);
    // get_random_date_time, utils.py:701-709
    int64_t t = s.start_epoch;
    if (s.random_start) {
      const int days = min(363, (int)(((uint64_t)q0.x * 364ull) >> 32));
      const int secs = min(86399, (int)(((uint64_t)q0.y * 86400ull) >> 32));
      t += (int64_t)days * 86400 + secs;
    }
    const double phase = s.random_phase ? u01(q0.z, q0.w) * 24.0 : 0.0;  // ClusterHouses.__init__, :789-792
    const Calendar cal = calendar_time((uint32_t)t);
    const double time_day = cal.hour + cal.minute * (1.0 / 60.0);
    const double od = p.od_amplitude * sin(p.two_pi_over_24 * (time_day + (-6 + phase))) + p.od_bias +
                      p.temp_std * gauss01(q1.x, q1.y, q1.z, q1.w);  // :793, :1070-1081
    double mp = 0.0;
    for (int i = 0; i < N; ++i) mp += s_cap[i] / p.hvac_cop;  // sequential, id order (:796-802)
    p.t_epoch[e] = t;
    const_cast<double*>(p.phase)[e] = phase;
    p.od_temp[e] = od;
    // PowerGrid.__init__, :1116: ratio * range ** (U * 2 - 1); :1182-1184: perlin seed = random()
    const_cast<double*>(p.artificial_ratio)[e] = s.artificial_ratio * pow(s.artificial_ratio_range, u01(q2.x, q2.y) * 2.0 - 1.0);
    const_cast<double*>(p.max_power)[e] = mp;
    p.base_power[e] = 0.0;
    p.signal[e] = 0.0;
    p.cluster_power[e] = 0.0;
    if (p.solar_gain != nullptr) p.solar_gain[e] = 0.0;
    if (p.time_since_interp != nullptr) p.time_since_interp[e] = s.interp_update_period + 1;
    if (p.perlin_seed != nullptr) const_cast<double*>(p.perlin_seed)[e] = u01(q2.z, q2.w);
  }
}

cudaError_t launch_populate(const KernelParams& kp_in, const MdrPopulationSpec& spec, const uint8_t* env_mask, double* ua,
                            double* cm, double* ca, double* hm, double* cap, double* target, double* deadband,
                            int32_t* lockout_dur, int precision, uint64_t draw_index, cudaStream_t stream) {
  KernelParams kp = kp_in;
  kp.precision = precision;
  const size_t smem = (size_t)kp.N * sizeof(double);
  populate_kernel<<<kp.E, 128, smem, stream>>>(kp, spec, env_mask, ua, cm, ca, hm, cap, target, deadband, lockout_dur, draw_index);
  return cudaGetLastError();
}


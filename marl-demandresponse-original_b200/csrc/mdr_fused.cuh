// Fused multi-step ("deploy") kernel with on-device metric accumulators, and its launcher.
// Included by mdr_kernels.cu inside namespace mdr (after the shared device helpers); not a standalone translation unit.
#pragma once

// ----------------------------------------------------------------------------------------
// Fused multi-step kernel (the "deploy" loop, main-deploy.py:102-209): K consecutive env steps in ONE
// launch for configurations whose per-step inputs are all produced on the device (on-device action
// source, Philox noise, constant or -- kInterp -- interpolated base power, no observation consumer).  A CTA owns G whole envs for
// the entire run: house state and coefficients stay in registers, nothing but the final state, the
// last reward and the per-env metric accumulators ever goes back to HBM.
//   * the per-env part of a step (clock, outdoor temperature, noise, grid signal) does not depend on
//     the houses here, so one warp computes it for 32 STEPS AT ONCE -- lane = step, every Philox draw
//     is counter-based on (env, step) -- into a double-buffered record ring, one batch ahead;
//   * the house warps then run 32 steps per batch with one house-warp barrier per step (cluster power
//     and, with metrics, the per-step max temperature error cross the warps through shared memory);
//   * metrics (main-deploy.py:124-209, metrics.py:22-47) that are sums over houses AND steps are kept
//     per thread in fp64 and reduced once at the end; only the per-step max needs the per-step exchange.
// Same arithmetic and the same Philox counters as K launches of the per-step kernels.
// ----------------------------------------------------------------------------------------
struct StepRec { double od_new, sig_new, gain; };

// one (env, step) record; mirrors env_prologue for base_power_mode == constant and on-device draws
// PowerGrid.step signal shapes (:1257-1314) are linear in the base power except regular_steps: signal =
// min(ratio * base * factor, max_power).  With interpolated base power the record warp cannot know the base of a
// future step (it depends on the houses at the refresh), so it hands over the factor instead.
__device__ __forceinline__ double signal_factor(const KernelParams& p, int time_sec, double noise) {
  const double two_pi = 2.0 * 3.141592653589793;
  if (p.signal_mode == MDR_SIG_SINUSOIDALS) {
    double f = 1.0;
    for (int i = 0; i < p.n_sinusoids; ++i) f += p.sin_ratios[i] * sin(mul_rn(two_pi, (double)time_sec) / p.sin_periods[i]);
    return f;
  }
  if (p.signal_mode == MDR_SIG_PERLIN) return fmax(0.0, 1.0 + p.perlin_amplitude * noise);
  return 1.0;  // flat
}

__device__ __noinline__ StepRec env_record(const KernelParams& p, int e2, uint32_t t, uint64_t step_index, bool factor_only) {
  Calendar cal = calendar_time(t);
  if (p.solar) calendar_date(cal);
  const bool perlin = p.signal_mode == MDR_SIG_PERLIN;
  double sig_noise = 0.0;
  if (perlin) {  // utils.Perlin.calculate_noise (utils.py:1247-1253) with hashed lattice gradients
    const int nb = p.perlin_nb_octaves;
    const double x = (double)cal.sod * p.inv_perlin_period;
    const uint64_t pkey = p.seed ^ (uint64_t)__double_as_longlong(p.perlin_seed[e2]);
    // (the 16-lane butterfly of env_prologue sums the same per-octave values in another order: ~1 ulp in fp64)
    for (int j = 0; j < nb; ++j) sig_noise += (double)perlin_octave(x, j, nb, p.perlin_octaves_step, pkey);
  }
  const double od_noise =
      p.temp_std * normal_from(philox4x32((uint32_t)(e2 + p.env_base), (uint32_t)step_index, (uint32_t)(step_index >> 32), STREAM_OD, p.seed));
  // ClusterHouses.compute_OD_temp, :1070-1081
  const double time_day = cal.hour + cal.minute * (1.0 / 60.0);
  StepRec rec;
  rec.od_new = p.od_amplitude * sin(p.two_pi_over_24 * (time_day + (-6 + p.phase[e2]))) + p.od_bias;
  rec.od_new += od_noise;
  rec.gain = p.solar ? solar_gain(cal, p.window_area, p.shading_coeff) : 0.0;
  const int time_sec = cal.hour * 3600 + cal.minute * 60 + cal.second;
  rec.sig_new = factor_only ? signal_factor(p, time_sec, sig_noise)
                            : grid_signal(p, p.avg_power_per_hvac * p.N, time_sec, sig_noise, p.artificial_ratio[e2], p.max_power[e2]);
  return rec;
}

template <typename R>
__device__ __forceinline__ R segmented_max(R v, int key, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const R tv = __shfl_down_sync(0xffffffffu, v, o);
    const int tk = __shfl_down_sync(0xffffffffu, key, o);
    if (lane + o < 32 && tk == key) v = fmax(v, tv);
  }
  return v;
}

struct FusedSmem {
  size_t off_rec, off_part, off_pmax, off_red, off_interp, total;
};
inline FusedSmem fused_smem_layout(int real_bytes, int genvs, int part_stride, bool metrics, int hmax, bool interp) {
  FusedSmem L;
  size_t o = 0;
  L.off_rec = o;  o += align16((size_t)2 * genvs * 32 * 4 * real_bytes);
  L.off_part = o; o += align16((size_t)2 * genvs * part_stride * real_bytes);
  L.off_pmax = o; o += metrics ? align16((size_t)2 * genvs * part_stride * real_bytes) : 0;
  L.off_red = o;  o += metrics ? align16((size_t)genvs * part_stride * 4 * sizeof(double)) : 0;
  L.off_interp = o; o += interp ? align16(((size_t)hmax + 2 * genvs) * sizeof(double)) : 0;  // values | base | last factor
  L.total = o;
  return L;
}

template <typename R, int kMaxThreads, int kAct, bool kMetrics, bool kInterp>
__global__ void __launch_bounds__(kMaxThreads, kMaxThreads <= 256 ? (kInterp && kMetrics ? 2 : 3) : 1) run_fused_kernel(const __grid_constant__ KernelParams p) {
  using T2 = typename Vec<R>::T2;
  using T4 = typename Vec<R>::T4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, G = p.G;
  const int env0 = blockIdx.x * G;
  const int genvs = min(G, p.E - env0);
  const int K = p.n_fused;
  const int n_batches = (K + 31) >> 5;
  T4* s_rec = reinterpret_cast<T4*>(smem_raw + p.off_env);  // [2][G][32] records (od_new, sig_new, gain, -) in the working precision
  const int dt = p.dt;

  // lane = step of the batch; one env after the other.  The houses consume the records in the working precision
  // (one 16-byte load per step); the fp64 values of the run's LAST step go straight to the per-env state.
  auto produce = [&](int b) {
    const int j = (b << 5) + lane;
    for (int le2 = 0; le2 < genvs; ++le2) {
      const int e2 = env0 + le2;
      const uint32_t t0 = (uint32_t)p.t_epoch[e2];
      if (j < K) {
        const StepRec rec = env_record(p, e2, t0 + (uint32_t)(j + 1) * (uint32_t)dt, step_now(p) + (uint64_t)j, kInterp);
        s_rec[((b & 1) * G + le2) * 32 + lane] = make4((R)rec.od_new, (R)rec.sig_new, (R)rec.gain, (R)0);
        if (j == K - 1) {
          p.od_temp[e2] = rec.od_new;
          if (kInterp) reinterpret_cast<double*>(smem_raw + p.off_in)[p.hmax + G + le2] = rec.sig_new;  // last factor, fp64
          else p.signal[e2] = rec.sig_new;
          if (p.solar) p.solar_gain[e2] = rec.gain;
        }
      }
    }
  };

  if (warp == p.pro_warp) {  // always a dedicated warp here (launch_fused refuses geometries without one)
    cta_sync();  // the house threads have read the initial od_temp / signal, which the last record overwrites
    for (int b = 0; b < n_batches; ++b) {
      produce(b);
      cta_sync();  // batch b ready; the house warps have finished batch b-1 (its buffer is free for b+1)
    }
    return;
  }

  const int H = genvs * N;
  const bool active = tid < H;
  const int le = active ? (N == 1 ? tid : (int)__umulhi((unsigned)tid, p.div_magic)) : 0;
  const int li = tid - le * N;
  const int e = env0 + le;
  const unsigned h = (unsigned)env0 * (unsigned)N + (unsigned)tid;
  R* s_part = reinterpret_cast<R*>(smem_raw + p.off_pw);    // [2][G][part_stride]
  R* s_pmax = reinterpret_cast<R*>(smem_raw + p.off_pen);   // [2][G][part_stride] (metrics only)
  const int first_warp = (le * N) >> 5;
  const int my_part = le * p.part_stride + (warp - first_warp);
  const int nparts = ((le * N + N - 1) >> 5) - first_warp + 1;
  const int part_buf = G * p.part_stride;
  const int part_row = le * p.part_stride;

  T2 tt = make2((R)0, (R)0);
  T4 ca4 = make4((R)0, (R)0, (R)0, (R)0), cb = ca4;
  T2 cc = make2((R)0, (R)1);
  int hv = 0;
  R od_old = 0, s_old = 0;
  if (active) {
    tt = reinterpret_cast<const T2*>(p.temps)[h];
    hv = p.hvac[h];
    ca4 = reinterpret_cast<const T4*>(p.coef_a)[h];
    cb = reinterpret_cast<const T4*>(p.coef_b)[h];
    cc = reinterpret_cast<const T2*>(p.coef_c)[h];
    od_old = (R)p.od_temp[e];
    s_old = (R)p.signal[e];
  }
  // interpolated base power (kInterp): per-env base / refresh clock live in registers of every house thread
  double base_cur = 0.0;
  R ratio_r = 0, maxp_r = 0;
  int tsi = 0, ikey = 0;
  double* s_ival = reinterpret_cast<double*>(smem_raw + p.off_in);  // [hmax] table values | [G] new base | [G] last factor
  if (kInterp && active) {
    base_cur = p.base_power[e];
    ratio_r = (R)p.artificial_ratio[e];
    maxp_r = (R)p.max_power[e];
    tsi = p.time_since_interp[e];
    ikey = p.interp_key[h];
  }
  cta_sync();  // (see the record warp)
  const R target = cb.w, p_on = cb.z, deadband = cc.x;
  const int lockdur = (int)cc.y;
  int on = hv & 1, lock = (hv >> 1) & 1, sso = hv >> 2;
  const R hi = target + deadband / 2, lo = target - deadband / 2;
  double acc_r = 0.0, acc_off = 0.0, acc_abs = 0.0, acc_sq = 0.0;              // per house, over the steps
  double m_maxsq = 0.0, m_max = 0.0, m_od = 0.0, m_sig = 0.0, m_cons = 0.0;   // per env (first house thread)
  double m_doff = 0.0, m_dabs = 0.0, m_dsq = 0.0;
  // lanes of this warp that belong to the same env (loop invariant)
  const int key = active ? le : -1;
  const int prev_key = __shfl_up_sync(0xffffffffu, key, 1);  // (unconditionally: a full-mask shuffle must not sit behind &&)
  const bool head = active && (lane == 0 || prev_key != key);
  const unsigned seg_mask = __match_any_sync(0xffffffffu, key);
  // cluster power as an integer redux.sync when every P_on of this CTA is an integral number of watts (the default
  // capacity lists / COP): the same value as the fp32 shuffle tree (exact either way), a fraction of the instructions
  const unsigned ip_on = (unsigned)p_on;
  int all_int;
  {
    const int mine = !active || (sizeof(R) == 4 && p_on >= (R)0 && p_on < (R)4194304 && (R)ip_on == p_on);
    asm volatile("{ .reg .pred a, b; setp.ne.s32 a, %1, 0; bar.red.and.pred b, 1, %2, a; selp.s32 %0, 1, 0, b; }"
                 : "=r"(all_int)
                 : "r"(mine), "r"(p.house_threads)
                 : "memory");
  }
  R b_r = 0, b_off = 0, b_abs = 0, b_sq = 0;                                   // ... of the current 32-step batch
  R e_maxsq = 0, e_max = 0, e_od = 0, e_sig = 0, e_cons = 0, e_doff = 0, e_dabs = 0, e_dsq = 0;
  R P = 0, reward = 0;

  for (int b = 0; b < n_batches; ++b) {
    cta_sync();  // batch b ready
    const int steps = min(32, K - (b << 5));
    const T4* recs = s_rec + ((b & 1) * G + le) * 32;
    for (int s = 0; s < steps; ++s) {
      const int j = (b << 5) + s;
      const T4 rec = recs[s];  // (od_new, sig_new, gain, -)
      R pw = 0, pen = 0, aerr = 0;
      if (active) {
        int cmd;
        if (kAct == MDR_ACT_BANGBANG) cmd = tt.x > target;  // agents/bangbang_controllers.py:50-61
        else {
          const uint64_t si = step_now(p) + (uint64_t)j;
          cmd = philox4x32(h + p.house_base, (uint32_t)si, (uint32_t)(si >> 32), STREAM_ACT, p.seed).x & 1;
        }
        // HVAC.step, :475-492
        if (!on) sso += dt;
        lock = !(on || sso >= lockdur);
        const int new_on = lock ? 0 : cmd;
        if (!lock && new_on) sso = 0;
        if (!lock && !new_on && sso + dt < lockdur) lock = 1;
        on = new_on;
        // SingleHouse.update_temperature, :681-738, with the OLD outdoor temperature and this step's solar gain
        const R qa = (on ? cb.y : (R)0) + (p.solar ? rec.z : (R)0);
        const R tss = od_old + qa * cb.x;
        const R x = tt.x - tss, y = tt.y - tss;
        tt.x = tt.x + (ca4.x * x + ca4.y * y);
        tt.y = tt.y + (ca4.z * x + ca4.w * y);
        pw = on ? p_on : (R)0;
        // utils.deadbandL2, utils.py:1266-1274
        if (hi < tt.x) pen = (tt.x - hi) * (tt.x - hi);
        else if (lo > tt.x) pen = (lo - tt.x) * (lo - tt.x);
        aerr = fabs(tt.x - target);
      }
      R* part = s_part + (j & 1) * part_buf;
      {
        R psum;
        if (all_int) psum = (R)__reduce_add_sync(seg_mask, on ? ip_on : 0u);
        else psum = segmented_sum<R>(pw, key, lane);
        if (head) part[my_part] = psum;
      }
      if (kMetrics) {
        R pm;
        if (sizeof(R) == 4) {
          // |err| >= 0: its bit pattern orders like an unsigned integer, so one redux.sync over the env's lanes does it
          pm = (R)__uint_as_float(__reduce_max_sync(seg_mask, __float_as_uint((float)aerr)));
        } else {
          pm = segmented_max<R>(aerr, key, lane);
        }
        if (head) s_pmax[(j & 1) * part_buf + my_part] = pm;
      }
      R sig_new = rec.y;
      if (!kInterp) {
        house_sync(p.house_threads);  // partials are double buffered by step parity: one rendezvous per step
      } else {
        // PowerGrid.step :1250-1255: the refresh clock advances, a due env re-interpolates its base power from the
        // houses' NEW state (N <= interp_nb_agents here: every house, summed in id order like :1218-1232)
        tsi += dt;
        const bool due = active && tsi >= p.interp_update_period;
        if (due) tsi = 0;
        int any_due;
        asm volatile("{ .reg .pred a, b; setp.ne.s32 a, %1, 0; bar.red.or.pred b, 1, %2, a; selp.s32 %0, 1, 0, b; }"
                     : "=r"(any_due)
                     : "r"((int)due), "r"(p.house_threads)
                     : "memory");
        if (any_due) {
          if (due) {
            const double tg = (double)target;
            s_ival[tid] = interp_eval<R>(p, ikey, (double)tt.x - tg, (double)tt.y - tg, (double)rec.x - tg, 0.0, 0.0);
          }
          house_sync(p.house_threads);
          if (due && li == 0) {
            double bsum = 0.0;
            for (int i = 0; i < N; ++i) bsum = add_rn(bsum, s_ival[le * N + i]);
            s_ival[p.hmax + le] = bsum;
          }
          house_sync(p.house_threads);
          if (due) base_cur = s_ival[p.hmax + le];
        }
        if (sizeof(R) == 4) sig_new = fminf((float)ratio_r * ((float)base_cur * (float)rec.y), (float)maxp_r);
        else sig_new = (R)fmin((double)ratio_r * (base_cur * (double)rec.y), (double)maxp_r);
      }
      if (active) {
        // fp32 mode keeps the per-step arithmetic of the pipelined kernel (fp32 power sums are exact: integer-valued
        // watts), fp64 mode that of the generic kernel; reg_signal_penalty :244-247 with the OLD signal, weighting :364-372
        R Ps = 0;
        if (nparts <= 8) {
#pragma unroll
          for (int w = 0; w < 8; ++w)
            if (w < nparts) Ps += part[part_row + w];
        } else {
          for (int w = 0; w < nparts; ++w) Ps += part[part_row + w];
        }
        P = Ps;
        R rew_r;
        if (sizeof(R) == 4) {
          const float dn = ((float)Ps - (float)s_old) * p.f_inv_n;
          rew_r = -((float)pen * p.f_k_temp + dn * dn * p.f_k_sig);
        } else {
          const double dn = ((double)Ps - (double)s_old) * p.inv_n;
          rew_r = (R)(-((double)pen * p.k_temp + dn * dn * p.k_sig));
        }
        reward = rew_r;
        if (kMetrics) {
          // batch-local accumulators in the working precision, flushed into fp64 every 32 steps
          const R err = tt.x - target;
          b_r += rew_r;
          b_off += err;
          b_abs += aerr;
          b_sq += err * err;
          if (li == 0) {
            R mx = 0;
            const R* pmx = s_pmax + (j & 1) * part_buf + part_row;
            if (nparts <= 8) {
#pragma unroll
              for (int w = 0; w < 8; ++w)
                if (w < nparts) mx = fmax(mx, pmx[w]);
            } else {
              for (int w = 0; w < nparts; ++w) mx = fmax(mx, pmx[w]);
            }
            const R sig = sig_new, d = sig - Ps;
            e_maxsq += mx * mx;
            e_max = fmax(e_max, mx);
            e_od += rec.x;
            e_sig += sig;
            e_cons += Ps;
            e_doff += d;
            e_dabs += fabs(d);
            e_dsq += d * d;
          }
        }
      }
      od_old = rec.x;
      s_old = sig_new;
    }
    if (kMetrics) {
      acc_r += (double)b_r; acc_off += (double)b_off; acc_abs += (double)b_abs; acc_sq += (double)b_sq;
      b_r = b_off = b_abs = b_sq = 0;
      if (li == 0) {
        m_maxsq += (double)e_maxsq; m_max = fmax(m_max, (double)e_max); m_od += (double)e_od; m_sig += (double)e_sig;
        m_cons += (double)e_cons; m_doff += (double)e_doff; m_dabs += (double)e_dabs; m_dsq += (double)e_dsq;
        e_maxsq = e_od = e_sig = e_cons = e_doff = e_dabs = e_dsq = 0;
      }
    }
  }

  // ---------------- write-back ---------------------------------------------------------------
  if (active) {
    reinterpret_cast<T2*>(p.temps)[h] = tt;
    p.hvac[h] = (sso << 2) | (lock << 1) | on;
    if (p.reward != nullptr) reinterpret_cast<R*>(p.reward)[h] = reward;
    if (li == 0) {
      p.cluster_power[e] = (double)P;
      p.t_epoch[e] = p.t_epoch[e] + (int64_t)K * dt;
      if (kInterp) {
        p.base_power[e] = base_cur;
        p.time_since_interp[e] = tsi;
        // the run's last signal in fp64 from the fp64 factor the record warp left behind
        p.signal[e] = fmin(mul_rn(mul_rn(base_cur, s_ival[p.hmax + G + le]), p.artificial_ratio[e]), p.max_power[e]);
      } else {
        p.base_power[e] = p.avg_power_per_hvac * N;
      }
    }
  }
  if (kMetrics) {
    // per-env totals of the per-house accumulators: warp-segmented sums, then the env's first thread adds the
    // warp partials in warp order (deterministic)
    double* s_red = reinterpret_cast<double*>(smem_raw + p.off_val);  // [G][part_stride][4]
    const double v0 = segmented_sum<double>(acc_r, key, lane), v1 = segmented_sum<double>(acc_off, key, lane);
    const double v2 = segmented_sum<double>(acc_abs, key, lane), v3 = segmented_sum<double>(acc_sq, key, lane);
    if (head) {
      double* d = s_red + (size_t)my_part * 4;
      d[0] = v0; d[1] = v1; d[2] = v2; d[3] = v3;
    }
    house_sync(p.house_threads);
    if (active && li == 0) {
      double t[4] = {0.0, 0.0, 0.0, 0.0};
      for (int w = 0; w < nparts; ++w)
        for (int k = 0; k < 4; ++k) t[k] += s_red[(size_t)(le * p.part_stride + w) * 4 + k];
      double* m = p.metrics + (size_t)e * MDR_N_METRICS;
      m[MDR_M_STEPS] += (double)K;
      m[MDR_M_SUM_MEAN_REWARD] += t[0] * p.inv_n;
      m[MDR_M_SUM_MEAN_TEMP_OFFSET] += t[1] * p.inv_n;
      m[MDR_M_SUM_MEAN_TEMP_ERROR] += t[2] * p.inv_n;
      m[MDR_M_SUM_SQ_TEMP_ERROR] += t[3];
      m[MDR_M_SUM_SQ_MAX_TEMP_ERROR] += m_maxsq;
      m[MDR_M_MAX_TEMP_ERROR] = fmax(m[MDR_M_MAX_TEMP_ERROR], m_max);
      m[MDR_M_SUM_OD_TEMP] += m_od;
      m[MDR_M_SUM_SIGNAL] += m_sig;
      m[MDR_M_SUM_CONSUMPTION] += m_cons;
      m[MDR_M_SUM_SIGNAL_OFFSET] += m_doff;
      m[MDR_M_SUM_SIGNAL_ERROR] += m_dabs;
      m[MDR_M_SUM_SQ_SIGNAL_ERROR] += m_dsq;
    }
  }
}

template <typename R, int kMaxThreads, int kAct, bool kMetrics, bool kInterp>
static cudaError_t launch_fused_k(const KernelParams& kp, const Geometry& g, size_t smem, cudaStream_t stream) {
  static std::atomic<uint64_t> latch{0};
  cudaError_t err = ensure_max_smem(run_fused_kernel<R, kMaxThreads, kAct, kMetrics, kInterp>, latch);
  if (err != cudaSuccess) return err;
  run_fused_kernel<R, kMaxThreads, kAct, kMetrics, kInterp><<<g.ctas, g.threads, smem, stream>>>(kp);
  return cudaGetLastError();
}

template <typename R, int kMaxThreads, int kAct>
static cudaError_t launch_fused_m(const KernelParams& kp, const Geometry& g, size_t smem, cudaStream_t stream) {
  const bool interp = kp.base_power_mode == MDR_BASE_INTERPOLATION;
  if (kp.metrics != nullptr)
    return interp ? launch_fused_k<R, kMaxThreads, kAct, true, true>(kp, g, smem, stream)
                  : launch_fused_k<R, kMaxThreads, kAct, true, false>(kp, g, smem, stream);
  return interp ? launch_fused_k<R, kMaxThreads, kAct, false, true>(kp, g, smem, stream)
                : launch_fused_k<R, kMaxThreads, kAct, false, false>(kp, g, smem, stream);
}

template <typename R, int kAct>
static cudaError_t launch_fused_t(const KernelParams& kp, const Geometry& g, size_t smem, cudaStream_t stream) {
  if (g.threads <= 256) return launch_fused_m<R, 256, kAct>(kp, g, smem, stream);
  if (g.threads <= 512) return launch_fused_m<R, 512, kAct>(kp, g, smem, stream);
  return launch_fused_m<R, 1024, kAct>(kp, g, smem, stream);
}

// plain steps that need nothing from the host between them (see run_fused_kernel)
bool fused_eligible(const KernelParams& kp) {
  // interpolated base power: every house is sampled (N <= interp_nb_agents), no solar gain (hour/date of the
  // interpolation point are 0 then), and a signal shape that is linear in the base power
  const bool base_ok = kp.base_power_mode == MDR_BASE_CONSTANT ||
                       (kp.N <= kp.interp_nb_agents && !kp.solar && kp.signal_mode != MDR_SIG_REGULAR_STEPS &&
                        kp.interp_table != nullptr && kp.interp_key != nullptr);
  return kp.is_reset == 0 && kp.obs == nullptr && (kp.action_source == MDR_ACT_BANGBANG || kp.action_source == MDR_ACT_RANDOM) &&
         base_ok &&
         kp.temp_penalty_mode == MDR_PEN_INDIVIDUAL_L2 && kp.od_noise == nullptr && kp.signal_noise == nullptr &&
         (kp.signal_mode != MDR_SIG_PERLIN || kp.perlin_seed != nullptr);
}

cudaError_t launch_fused(const KernelParams& kp_in, const Geometry& g, int precision, int n_steps, cudaStream_t stream) {
  if (g.pro_warp < g.house_warps) return cudaErrorInvalidConfiguration;  // needs the dedicated record warp (N <= 992)
  KernelParams kp = kp_in;
  kp.n_fused = n_steps;
  const int rb = precision;
  const FusedSmem L = fused_smem_layout(rb, g.envs_per_cta, g.part_stride, kp.metrics != nullptr, g.hmax,
                                        kp.base_power_mode == MDR_BASE_INTERPOLATION);
  kp.off_env = (int)L.off_rec; kp.off_pw = (int)L.off_part; kp.off_pen = (int)L.off_pmax; kp.off_val = (int)L.off_red;
  kp.off_in = (int)L.off_interp;
  if (precision == MDR_F32)
    return kp.action_source == MDR_ACT_BANGBANG ? launch_fused_t<float, MDR_ACT_BANGBANG>(kp, g, L.total, stream)
                                                 : launch_fused_t<float, MDR_ACT_RANDOM>(kp, g, L.total, stream);
  return kp.action_source == MDR_ACT_BANGBANG ? launch_fused_t<double, MDR_ACT_BANGBANG>(kp, g, L.total, stream)
                                               : launch_fused_t<double, MDR_ACT_RANDOM>(kp, g, L.total, stream);
}


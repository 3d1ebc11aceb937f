// One CTA walks ONE whole env of any size in sub-tiles -- the step without an observation (deploy loops, Monte-Carlo
// runs, VecDemandResponseEnv(with_obs=False)) for clusters beyond the pipelined kernel's 224 houses per tile.
// Included by mdr_kernels.cu inside namespace mdr; not a standalone translation unit.
//
// Without observation rows nothing of a house's step depends on its neighbours: the only cluster-wide quantity is the
// power sum the reward needs (compute_rewards :330-373 with reg_signal_penalty :244-247), and that can wait until the
// CTA has seen every house.  So, instead of splitting the env over a thread-block cluster that meets at every tile
// (mdr_pipe_split.cuh: 35 us per step on 1 000 x 1 000, the DSMEM rendezvous per 200 houses is what it costs), one
// CTA keeps the env to itself:
//   pass 1  warps 0-6, 224 houses at a time: load, lockout machine (HVAC.step :463-492), affine ETP update
//           (update_temperature :664-738), state store; the penalty goes to shared memory, the power into a per-thread sum;
//           warp 7 evaluates the env's prologue meanwhile (clock, outdoor temperature, noise, signal; env_prologue);
//   --      one CTA barrier: cluster power from the warp partials in warp order (deterministic; exact for integer watts);
//   [refresh, 1 step in 75: table walk on the sampled houses' NEW state, sum in id order like the generic kernel]
//   pass 2  all warps: reward of every house from its parked penalty and the env's power; per-env outputs by thread 0.
// 1 000 envs x 1 000 houses are 1 000 independent CTAs, a wave and a half of the GPU with no inter-CTA traffic at all.
// Same arithmetic, operation for operation, as the generic kernel (mdr::step_kernel) -- the tests compare them bitwise.
#pragma once

constexpr int kWideThreads = 256;
constexpr int kWideHouseThreads = 224;  // warps 0-6 own houses in pass 1, warp 7 the env

struct WideSmem {
  EnvScratch env;
  double part[8];
};

template <typename R, int kAct>
__global__ void __launch_bounds__(kWideThreads, 4) step_wide_kernel(const __grid_constant__ KernelParams p) {
  using T2 = typename Vec<R>::T2;
  using T4 = typename Vec<R>::T4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WideSmem& sm = *reinterpret_cast<WideSmem*>(smem_raw);
  double* const s_val = reinterpret_cast<double*>(smem_raw + p.off_val);  // [sampled houses] refresh values
  R* const s_pen = reinterpret_cast<R*>(smem_raw + p.off_pen);             // [N] deadband penalties
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  const int e = blockIdx.x;
  const size_t h0 = (size_t)e * N;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything the previous launch wrote is visible from here on

  if (warp == 7) {  // 16 lanes share the env's draws (perlin octaves + the outdoor-temperature normal)
    PipeEnv unused;
    env_prologue<false>(p, sm.env, unused, e, lane & 15, 16, lane < 16, false, false);
  }
  if (p.solar) cta_sync();  // the thermal update needs this step's solar gain (evaluated at the NEW datetime, :694)

  double psum = 0.0;
  if (warp < 7) {
    const R od_old = (R)p.od_temp[e];  // (the env's outputs are written at the very end)
    const R gain = p.solar ? (R)sm.env.gain_now : (R)0;
    const int dt = p.dt;
    for (int i = tid; i < N; i += kWideHouseThreads) {
      const size_t h = h0 + i;
      T2 tt = reinterpret_cast<const T2*>(p.temps)[h];
      const int hv = p.hvac[h];
      const T4 cb = reinterpret_cast<const T4*>(p.coef_b)[h];
      const T2 cc = reinterpret_cast<const T2*>(p.coef_c)[h];
      const T4 ca4 = reinterpret_cast<const T4*>(p.coef_a)[h];
      const R target = cb.w, p_on = cb.z, deadband = cc.x;
      int cmd;
      if (kAct == MDR_ACT_ARRAY) cmd = p.actions[h] != 0;
      else if (kAct == MDR_ACT_BANGBANG) cmd = tt.x > target;  // agents/bangbang_controllers.py:50-61
      else cmd = philox4x32((uint32_t)h + p.house_base, (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32), STREAM_ACT, p.seed).x & 1;
      // HVAC.step, :475-492
      int on = hv & 1, sso = hv >> 2;
      const int lockdur = (int)cc.y;
      if (!on) sso += dt;
      int lock = !(on || sso >= lockdur);
      const int new_on = lock ? 0 : cmd;
      if (!lock && new_on) sso = 0;
      if (!lock && !new_on && sso + dt < lockdur) lock = 1;
      on = new_on;
      // SingleHouse.update_temperature, :681-738, with the OLD outdoor temperature
      const R qa = (on ? cb.y : (R)0) + gain;
      const R tss = od_old + qa * cb.x;
      const R x = tt.x - tss, y = tt.y - tss;
      tt.x = tt.x + (ca4.x * x + ca4.y * y);
      tt.y = tt.y + (ca4.z * x + ca4.w * y);
      reinterpret_cast<T2*>(p.temps)[h] = tt;
      p.hvac[h] = (sso << 2) | (lock << 1) | on;
      psum += (double)(on ? p_on : (R)0);
      // utils.deadbandL2, utils.py:1266-1274
      R pen = 0;
      const R hi = target + deadband / 2, lo = target - deadband / 2;
      if (hi < tt.x) pen = (tt.x - hi) * (tt.x - hi);
      else if (lo > tt.x) pen = (lo - tt.x) * (lo - tt.x);
      s_pen[i] = pen;
    }
    psum = warp_sum(psum);
    if (lane == 0) sm.part[warp] = psum;
  }
  cta_sync();  // the env's record, the warp partials, the parked penalties and the new state (for the refresh) are complete
  double P = 0.0;
#pragma unroll
  for (int w = 0; w < 7; ++w) P += sm.part[w];

  // interpolation refresh (PowerGrid.step :1250-1255, interpolatePower :1195-1234), CTA-uniform
  if (p.base_power_mode == MDR_BASE_INTERPOLATION && sm.env.due) {
    const int nb = p.interp_nb_agents;
    const int nsamp = N <= nb ? N : nb;
    for (int li = tid; li < nsamp; li += kWideThreads) {
      int src = li;
      if (N > nb) {
        if (p.interp_ids) src = p.interp_ids[(size_t)e * nb + li];
        else {
          const uint4 r = philox4x32((uint32_t)(e + p.env_base), (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32),
                                     STREAM_IDS + 16 * (uint32_t)li, p.seed);
          src = (int)(((uint64_t)r.x * (uint64_t)N) >> 32);
        }
      }
      const size_t hs = h0 + src;
      const T2 t2 = reinterpret_cast<const T2*>(p.temps)[hs];
      const double tg = (double)reinterpret_cast<const T4*>(p.coef_b)[hs].w;
      s_val[li] = interp_eval<R>(p, p.interp_key[hs], (double)t2.x - tg, (double)t2.y - tg, sm.env.od_new - tg, sm.env.hour_s,
                                 sm.env.date);
    }
    cta_sync();
    if (tid == 0) {
      double base = 0.0;
      for (int i = 0; i < nsamp; ++i) base = add_rn(base, s_val[i]);  // id order, :1218-1232
      if (N > nb) base = mul_rn(base, (double)N / (double)nb);
      const double sig = grid_signal(p, base, sm.env.time_sec, sm.env.sig_noise, p.artificial_ratio[e], p.max_power[e]);
      p.base_power[e] = base;
      p.time_since_interp[e] = 0;
      p.signal[e] = sig;
    }
  }

  // per-env state, by the env's first thread (like the generic kernel's env head)
  if (tid == 0) {
    const EnvScratch& es = sm.env;
    p.cluster_power[e] = P;
    p.od_temp[e] = es.od_new;
    p.t_epoch[e] = (int64_t)es.t_new;
    if (p.solar) p.solar_gain[e] = es.gain_now;
    if (!es.due) {
      p.base_power[e] = es.base;
      p.signal[e] = es.sig_new;
      if (p.base_power_mode == MDR_BASE_INTERPOLATION) p.time_since_interp[e] = es.tsi;
    }
  }

  // pass 2: reward (reg_signal_penalty :244-247 with the OLD signal; weighting :364-372)
  if (p.reward != nullptr) {
    const double dn = (P - sm.env.s_old) * p.inv_n;
    const double sig_term = dn * dn * p.k_sig;
    R* const rew = reinterpret_cast<R*>(p.reward) + h0;
    for (int i = tid; i < N; i += kWideThreads) rew[i] = (R)(-((double)s_pen[i] * p.k_temp + sig_term));
  }
}

// the step without an observation on clusters of 225 .. 8 192 houses (individual_L2 penalty, no greedy controller, no metric
// accumulators, no masked reset); needs enough envs to fill the GPU with one CTA each
bool wide_eligible(const KernelParams& kp) {
  return kp.obs == nullptr && kp.is_reset == 0 && kp.N > 224 && kp.N <= 8192 && kp.E >= 64 &&
         kp.temp_penalty_mode == MDR_PEN_INDIVIDUAL_L2 && kp.action_source != MDR_ACT_GREEDY && kp.metrics == nullptr &&
         kp.env_mask == nullptr;
}

template <typename R, int kAct>
static cudaError_t launch_wide_t(const KernelParams& kp_in, bool no_pdl, cudaStream_t stream) {
  static std::atomic<uint64_t> latch{0};
  cudaError_t err = ensure_max_smem(step_wide_kernel<R, kAct>, latch);
  if (err != cudaSuccess) return err;
  KernelParams kp = kp_in;
  size_t o = align16(sizeof(WideSmem));
  kp.off_val = (int)o;
  if (kp.base_power_mode == MDR_BASE_INTERPOLATION) o += align16((size_t)(kp.N <= kp.interp_nb_agents ? kp.N : kp.interp_nb_agents) * sizeof(double));
  kp.off_pen = (int)o;
  o += align16((size_t)kp.N * sizeof(R));
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(kp.E);
  lc.blockDim = dim3(kWideThreads);
  lc.dynamicSmemBytes = o;
  lc.stream = stream;
  lc.attrs = &attr;
  lc.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&lc, step_wide_kernel<R, kAct>, kp);
}

template <typename R>
static cudaError_t launch_wide_r(const KernelParams& kp, bool no_pdl, cudaStream_t stream) {
  if (kp.action_source == MDR_ACT_ARRAY) return launch_wide_t<R, MDR_ACT_ARRAY>(kp, no_pdl, stream);
  if (kp.action_source == MDR_ACT_BANGBANG) return launch_wide_t<R, MDR_ACT_BANGBANG>(kp, no_pdl, stream);
  return launch_wide_t<R, MDR_ACT_RANDOM>(kp, no_pdl, stream);
}

cudaError_t launch_wide(const KernelParams& kp, int precision, bool no_pdl, cudaStream_t stream) {
  return precision == MDR_F32 ? launch_wide_r<float>(kp, no_pdl, stream) : launch_wide_r<double>(kp, no_pdl, stream);
}

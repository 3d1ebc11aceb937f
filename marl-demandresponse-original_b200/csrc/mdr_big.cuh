// Envs larger than a thread-block cluster (N > 16 384 houses; SURVEY section 7 step 6 "two-pass beyond"): plain CTAs,
// three launches per step, the cluster-wide quantities cross the CTAs through a small global workspace.
// Included by mdr_kernels.cu inside namespace mdr; not a standalone translation unit.
//   1. big_update_kernel : per house -- lockout machine, affine ETP update, state store; per-CTA totals of power and
//                          penalties (sum of v/N, max) into the workspace.  Reads the per-env state, never writes it.
//   2. big_env_kernel    : one CTA per env -- the per-env prologue (clock, outdoor temperature, noise, signal), the CTA
//                          totals added in CTA order (deterministic), the interpolation refresh on the houses' NEW state,
//                          per-env outputs; leaves an EnvScratch record per env in the workspace.
//   3. big_finish_kernel : per house -- reward and the observation row (any flag / neighbour mode; messages are
//                          recomputed from the neighbours' state in global memory).
// ClusterHouses.step :1005-1055, compute_rewards :330-373, PowerGrid.step :1236-1316, make_cluster_obs_dict :904-1003.
#pragma once

constexpr int kBigThreads = 256;

struct BigPart { double P, pen_sum, pen_max, pad; };

inline size_t big_workspace_bytes(int n_envs, int n_houses) {
  const size_t nparts = ((size_t)n_houses + kBigThreads - 1) / kBigThreads;
  return align16((size_t)n_envs * sizeof(EnvScratch)) + (size_t)n_envs * nparts * sizeof(BigPart);
}

template <typename R>
__global__ void __launch_bounds__(kBigThreads) big_update_kernel(const __grid_constant__ KernelParams p, BigPart* parts, int nparts) {
  using T2 = typename Vec<R>::T2;
  using T4 = typename Vec<R>::T4;
  __shared__ double s_red[3][kBigThreads / 32];
  __shared__ double s_gain;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e = blockIdx.x / nparts, part = blockIdx.x - e * nparts;
  const int N = p.N;
  const int li = part * kBigThreads + tid;
  const bool reset = p.is_reset != 0;
  const bool active = li < N && (p.env_mask == nullptr || p.env_mask[e] != 0);
  const size_t h = (size_t)e * N + li;
  if (tid == 0) {
    // SingleHouse.update_temperature evaluates house_solar_gain at the NEW datetime (:694)
    double gain = 0.0;
    if (p.solar && !reset) {
      Calendar cal = calendar_time((uint32_t)p.t_epoch[e] + (uint32_t)p.dt);
      calendar_date(cal);
      gain = solar_gain(cal, p.window_area, p.shading_coeff);
    }
    s_gain = gain;
  }
  __syncthreads();
  double pw = 0.0, pen = 0.0;
  if (active) {
    T2 tt = reinterpret_cast<const T2*>(p.temps)[h];
    const T4 cb = reinterpret_cast<const T4*>(p.coef_b)[h];
    const T2 cc = reinterpret_cast<const T2*>(p.coef_c)[h];
    int hv = p.hvac[h];
    int on = hv & 1, sso = hv >> 2;
    const R target = cb.w, deadband = cc.x;
    if (!reset) {
      const T4 ca4 = reinterpret_cast<const T4*>(p.coef_a)[h];
      int cmd;
      if (p.action_source == MDR_ACT_ARRAY) cmd = p.actions[h] != 0;
      else if (p.action_source == MDR_ACT_BANGBANG) cmd = tt.x > target;  // agents/bangbang_controllers.py:50-61
      else cmd = philox4x32((uint32_t)h, (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32), STREAM_ACT, p.seed).x & 1;
      // HVAC.step, :475-492
      const int dt = p.dt, lockdur = (int)cc.y;
      if (!on) sso += dt;
      int lock = !(on || sso >= lockdur);
      const int new_on = lock ? 0 : cmd;
      if (!lock && new_on) sso = 0;
      if (!lock && !new_on && sso + dt < lockdur) lock = 1;
      on = new_on;
      // SingleHouse.update_temperature, :681-738, with the OLD outdoor temperature
      const R od_old = (R)p.od_temp[e];
      const R qa = (on ? cb.y : (R)0) + (R)s_gain;
      const R tss = od_old + qa * cb.x;
      const R x = tt.x - tss, y = tt.y - tss;
      tt.x = tt.x + (ca4.x * x + ca4.y * y);
      tt.y = tt.y + (ca4.z * x + ca4.w * y);
      reinterpret_cast<T2*>(p.temps)[h] = tt;
      p.hvac[h] = (sso << 2) | (lock << 1) | on;
    }
    pw = on ? (double)cb.z : 0.0;
    // utils.deadbandL2, utils.py:1266-1274
    const R hi = target + deadband / 2, lo = target - deadband / 2;
    R pr = 0;
    if (hi < tt.x) pr = (tt.x - hi) * (tt.x - hi);
    else if (lo > tt.x) pr = (lo - tt.x) * (lo - tt.x);
    pen = (double)pr;
  }
  const double ws = warp_sum(pw), wn = warp_sum(pen / N), wm = warp_max(pen);
  if (lane == 0) { s_red[0][warp] = ws; s_red[1][warp] = wn; s_red[2][warp] = wm; }
  __syncthreads();
  if (tid == 0) {
    BigPart bp = {0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < kBigThreads / 32; ++w) {
      bp.P += s_red[0][w];
      bp.pen_sum += s_red[1][w];
      bp.pen_max = fmax(bp.pen_max, s_red[2][w]);
    }
    parts[(size_t)e * nparts + part] = bp;
  }
}

template <typename R>
__global__ void __launch_bounds__(128) big_env_kernel(const __grid_constant__ KernelParams p, EnvScratch* recs, const BigPart* parts,
                                                      int nparts) {
  using T2 = typename Vec<R>::T2;
  using T4 = typename Vec<R>::T4;
  __shared__ EnvScratch s_es;
  __shared__ double s_val[MDR_MAX_HOUSES_PER_ENV];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e = blockIdx.x;
  const int N = p.N;
  const bool reset = p.is_reset != 0, observe_only = p.is_reset == 2;
  if (p.env_mask != nullptr && p.env_mask[e] == 0) return;
  if (warp == 0) {
    PipeEnv unused;
    env_prologue<false>(p, s_es, unused, e, lane & 15, 16, lane < 16, reset, observe_only);
    // CTA totals in CTA order: lane-strided, then the fixed shuffle tree (deterministic)
    double ps = 0.0, pn = 0.0, pm = 0.0;
    for (int i = lane; i < nparts; i += 32) {
      const BigPart bp = parts[(size_t)e * nparts + i];
      ps += bp.P; pn += bp.pen_sum; pm = fmax(pm, bp.pen_max);
    }
    ps = warp_sum(ps); pn = warp_sum(pn); pm = warp_max(pm);
    __syncwarp();
    if (lane == 0) { s_es.P = ps; s_es.pen_mean = pn; s_es.pen_max = pm; }
  }
  __syncthreads();
  const bool interp_mode = p.base_power_mode == MDR_BASE_INTERPOLATION;
  if (tid == 0 && !observe_only) {
    p.cluster_power[e] = s_es.P;
    if (!reset) {
      p.od_temp[e] = s_es.od_new;
      p.t_epoch[e] = (int64_t)s_es.t_new;
    }
    if (p.solar) p.solar_gain[e] = s_es.gain_now;
    if (!s_es.due) {
      p.base_power[e] = s_es.base;
      p.signal[e] = s_es.sig_new;
      if (interp_mode) p.time_since_interp[e] = s_es.tsi;
    }
  }
  if (s_es.due) {  // PowerGrid.step :1250-1255, interpolatePower :1195-1234 on the houses' NEW state
    const int nb = p.interp_nb_agents;
    const int nsamp = N <= nb ? N : nb;
    for (int i = tid; i < nsamp; i += blockDim.x) {
      int src = i;
      if (N > nb) {
        if (p.interp_ids) src = p.interp_ids[(size_t)e * nb + i];
        else {
          const uint4 r = philox4x32((uint32_t)e, (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32),
                                     STREAM_IDS + 16 * (uint32_t)i, p.seed);
          src = (int)(((uint64_t)r.x * (uint64_t)N) >> 32);
        }
      }
      const size_t hs = (size_t)e * N + src;
      const T2 t2 = reinterpret_cast<const T2*>(p.temps)[hs];
      const double tg = (double)reinterpret_cast<const T4*>(p.coef_b)[hs].w;
      s_val[i] = interp_eval<R>(p, p.interp_key[hs], (double)t2.x - tg, (double)t2.y - tg, s_es.od_new - tg, s_es.hour_s, s_es.date);
    }
    __syncthreads();
    if (tid == 0) {
      double base = 0.0;
      for (int i = 0; i < nsamp; ++i) base = add_rn(base, s_val[i]);  // id order, :1218-1232
      if (N > nb) base = mul_rn(base, (double)N / (double)nb);
      const double sig = grid_signal(p, base, s_es.time_sec, s_es.sig_noise, p.artificial_ratio[e], p.max_power[e]);
      p.base_power[e] = base;
      p.time_since_interp[e] = 0;
      p.signal[e] = sig;
      s_es.f_sig = sig * p.inv_norm_sig_agents;
      s_es.sig_new = sig;
    }
    __syncthreads();
  }
  if (tid == 0) recs[e] = s_es;
}

template <typename R>
__global__ void __launch_bounds__(kBigThreads) big_finish_kernel(const __grid_constant__ KernelParams p, const EnvScratch* recs,
                                                                 int nparts) {
  using T2 = typename Vec<R>::T2;
  using T4 = typename Vec<R>::T4;
  __shared__ EnvScratch s_es;
  const int tid = threadIdx.x;
  const int e = blockIdx.x / nparts, part = blockIdx.x - e * nparts;
  const int N = p.N;
  const int li = part * kBigThreads + tid;
  if (p.env_mask != nullptr && p.env_mask[e] == 0) return;
  if (tid == 0) s_es = recs[e];
  __syncthreads();
  if (li >= N) return;
  const bool reset = p.is_reset != 0;
  const size_t h = (size_t)e * N + li;
  const T2 tt = reinterpret_cast<const T2*>(p.temps)[h];
  const T4 cb = reinterpret_cast<const T4*>(p.coef_b)[h];
  const T2 cc = reinterpret_cast<const T2*>(p.coef_c)[h];
  const int hv = p.hvac[h];
  const R target = cb.w, deadband = cc.x;
  const double P = s_es.P;
  if (!reset && p.reward != nullptr) {
    const R hi = target + deadband / 2, lo = target - deadband / 2;
    R pen = 0;
    if (hi < tt.x) pen = (tt.x - hi) * (tt.x - hi);
    else if (lo > tt.x) pen = (lo - tt.x) * (lo - tt.x);
    double tp = (double)pen;
    const int pen_mode = p.temp_penalty_mode;
    if (pen_mode == MDR_PEN_COMMON_L2) tp = s_es.pen_mean;
    else if (pen_mode == MDR_PEN_COMMON_MAX) tp = s_es.pen_max;
    else if (pen_mode == MDR_PEN_MIXTURE)
      tp = (p.mix_alpha_ind * tp + p.mix_alpha_common * s_es.pen_mean + p.mix_alpha_max * s_es.pen_max) /
           (p.mix_alpha_ind + p.mix_alpha_common + p.mix_alpha_max);
    // reg_signal_penalty :244-247 with the OLD signal; weighting :364-372
    const double dn = (P - s_es.s_old) * p.inv_n;
    reinterpret_cast<R*>(p.reward)[h] = (R)(-(tp * p.k_temp + dn * dn * p.k_sig));
  }
  if (p.obs == nullptr) return;
  const R inv_norm = (R)p.inv_norm_reg_sig;
  // SingleHouse.message :624-662 of house j, from its (new) state in global memory
  auto msg_at = [&](int j) -> T4 {
    const size_t hj = (size_t)e * N + j;
    const T2 tj = reinterpret_cast<const T2*>(p.temps)[hj];
    const T4 cj = reinterpret_cast<const T4*>(p.coef_b)[hj];
    const int hvj = p.hvac[hj];
    return make4((tj.x - cj.w) * (R)0.2, (R)(hvj >> 2), ((hvj & 1) ? cj.z : (R)0) * inv_norm, cj.z * inv_norm);
  };
  HouseRow<R> hr;
  hr.t_air = tt.x; hr.t_mass = tt.y; hr.target = target; hr.deadband = deadband; hr.p_on = cb.z; hr.inv_lock = inv_real(cc.y);
  hr.on = hv & 1; hr.lock = (hv >> 1) & 1; hr.sso = hv >> 2; hr.P = P; hr.h = (unsigned)h; hr.e = e; hr.li = li;
  generic_row<R>(p, reinterpret_cast<R*>(p.obs) + h * p.F, hr, s_es, p.state_flags, p.msg_flags, p.comm_mode,
                 p.msg_keep != nullptr, p.comm_defect_prob > 0.0, p.C, msg_at);
}

template <typename R>
static cudaError_t launch_big_t(const KernelParams& kp, void* workspace, cudaStream_t stream) {
  const int nparts = (kp.N + kBigThreads - 1) / kBigThreads;
  EnvScratch* recs = reinterpret_cast<EnvScratch*>(workspace);
  BigPart* parts = reinterpret_cast<BigPart*>(reinterpret_cast<unsigned char*>(workspace) + align16((size_t)kp.E * sizeof(EnvScratch)));
  const unsigned grid = (unsigned)kp.E * (unsigned)nparts;
  big_update_kernel<R><<<grid, kBigThreads, 0, stream>>>(kp, parts, nparts);
  big_env_kernel<R><<<(unsigned)kp.E, 128, 0, stream>>>(kp, recs, parts, nparts);
  if (kp.obs != nullptr || (kp.reward != nullptr && kp.is_reset == 0))
    big_finish_kernel<R><<<grid, kBigThreads, 0, stream>>>(kp, recs, nparts);
  return cudaGetLastError();
}

cudaError_t launch_big(const KernelParams& kp, int precision, void* workspace, cudaStream_t stream) {
  return precision == MDR_F32 ? launch_big_t<float>(kp, workspace, stream) : launch_big_t<double>(kp, workspace, stream);
}

size_t big_workspace(int n_envs, int n_houses) { return big_workspace_bytes(n_envs, n_houses); }

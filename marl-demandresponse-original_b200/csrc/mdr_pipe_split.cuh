// Persistent software-pipelined step kernel for envs of 225 .. 1792 houses (fp32, default observation layout):
// the env is split over the `cl` CTAs of a thread-block cluster, `cl_slice` houses each, and the persistent grid is a
// multiple of `cl` CTAs, so tile t = blockIdx.x + it * gridDim.x is the slice of rank t % cl of env t / cl and the cl
// tiles of an env are always held by the cl CTAs of one cluster in the same iteration.
// Same pipeline as mdr_pipe.cuh (cp.async input stages, prologue warp + mbarrier ring, bulk observation store); what the
// env's CTAs owe each other crosses distributed shared memory, pushed by the producer:
//   * every warp's partial power sum (and metric partials) into all cl CTAs' [rank][warp] slots,
//   * the first / last houses' messages into the neighbouring CTAs' halo entries of the message window,
// as asynchronous remote stores that credit their bytes to the DESTINATION CTA's mbarrier (st.async ... complete_tx); the
// tile's only rendezvous is the wait on the CTA's own mbarrier (its own warps arrive, one of them announcing the bytes
// the peers owe) -- it replaces the CTA barrier of the single-CTA kernel; no fence, no remote arrive.  Windows, partial
// slots and mbarriers are double buffered by tile parity.
// The prologue warp of every CTA evaluates the env's record for itself (same inputs, same result); the per-env outputs
// are written once, by rank 0, after the rendezvous (nobody may overwrite what a peer's prologue can still read).
// An interpolation refresh due for the env is evaluated inside the tile (by every CTA, redundantly: it is rare).
// Included by mdr_kernels.cu inside namespace mdr; not a standalone translation unit.
#pragma once

constexpr int kMaxSplit = 8;  // portable cluster size; 8 x 224 houses (fp32 power sums stay exact below 2^24 W)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Asynchronous remote stores (DSMEM) that signal the DESTINATION CTA's mbarrier with the bytes they deliver
// (complete_tx): the consumer's wait on its own mbarrier is all the synchronisation the data needs -- no release fence,
// no remote arrive on the producer's side (a fence + 5 arrives per warp cost 1-3 us per tile).
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(raddr), "f"(v), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, const float4 v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// One release fence at cluster scope, then relaxed arrives on every peer: a release per arrive would pay the fence
// (and wait for everything this thread has in flight) once per peer -- measured 2.8 us per tile with 5 peers.
__device__ __forceinline__ void fence_release_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, int parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT_CL:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_CL;\n"
      "bra LAB_WAIT_CL;\n"
      "DONE_CL:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct SplitCtl {
  PipeCtl pipe;      // prologue ring (first member: prologue_pass addresses it through off_ctl)
  uint64_t xbar[2];  // the env's rendezvous, by tile parity (count cl x house warps)
  uint64_t rbar;     // second rendezvous of a tile with an interpolation refresh: the peers' NEW state is in global memory
};

// interpolation refresh of a split env inside the tile loop (PowerGrid.step :1250-1255, interpolatePower :1195-1234):
// every CTA of the cluster evaluates the sampled houses (their NEW temperatures are visible since the rendezvous)
// and sums them in id order -- identical results in every CTA, no second exchange
__device__ __noinline__ double split_refresh(const KernelParams& p, int e, const PipeEnv& pe, int T, int n_refresh) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_val = reinterpret_cast<double*>(smem_raw + p.off_val);  // [T] values, then [T] = the new signal
  const int tid = threadIdx.x, N = p.N, nb = p.interp_nb_agents;
  {
    // the tile's state stores were issued after its rendezvous: one more, so that every CTA's new temperatures are
    // visible (release / acquire at cluster scope covers global memory) before anybody samples them
    SplitCtl& sctl = *reinterpret_cast<SplitCtl*>(smem_raw + p.off_ctl);
    __syncwarp();
    if ((tid & 31) == 0) {
      fence_release_cluster();
      const uint32_t a = smem_u32(&sctl.rbar);
      for (int r = 0; r < p.cl; ++r) mbar_arrive_remote(mapa_shared(a, r));
    }
    mbar_wait_cluster(&sctl.rbar, n_refresh & 1);
  }
  const int nsamp = N <= nb ? N : nb;
  double hour_s = 0.0, date = 0.0;
  if (p.solar) {
    Calendar cal = calendar_time(pe.t_new);
    calendar_date(cal);
    hour_s = (double)cal.sod;
    date = (double)cal.yday;
  }
  if (tid < nsamp) {
    int src = tid;
    if (N > nb) {
      if (p.interp_ids) src = p.interp_ids[(size_t)e * nb + tid];
      else {
        const uint4 r = philox4x32((uint32_t)(e + p.env_base), (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32),
                                   STREAM_IDS + 16 * (uint32_t)tid, p.seed);
        src = (int)(((uint64_t)r.x * (uint64_t)N) >> 32);
      }
    }
    const size_t hs = (size_t)e * N + src;
    const float2 t2 = reinterpret_cast<const float2*>(p.temps)[hs];
    const double tg = (double)reinterpret_cast<const float4*>(p.coef_b)[hs].w;
    s_val[tid] = interp_eval_grid<float, InterpGrid>(*reinterpret_cast<const InterpGrid*>(smem_raw + p.off_grid), p.interp_table,
                                                     p.interp_key[hs], (double)t2.x - tg, (double)t2.y - tg, pe.od_new - tg, hour_s, date);
  }
  house_sync(T);
  if (tid == 0) {
    double base = 0.0;
    for (int i = 0; i < nsamp; ++i) base = add_rn(base, s_val[i]);  // id order, :1218-1232
    if (N > nb) base = mul_rn(base, (double)N / (double)nb);
    s_val[T] = base;
    s_val[T + 1] = grid_signal(p, base, pe.time_sec, pe.sig_noise, p.artificial_ratio[e], p.max_power[e]);
  }
  house_sync(T);
  return s_val[T + 1];
}

template <int kC, int kAct, bool kObs, bool kMetrics>
__global__ void __launch_bounds__(256, 3) step_pipe_split_kernel(const __grid_constant__ KernelParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  SplitCtl& sctl = *reinterpret_cast<SplitCtl*>(smem_raw + p.off_ctl);
  PipeCtl& ctl = sctl.pipe;
  const int ring_mask = 2 * p.pro_batch - 1;
  const int ring_shift = 31 - __clz(ring_mask + 1);
  const int ncl = p.cl;
  const int rank = (int)cluster_ctarank();
  MDR_CTA_STAMP(0);
  if (p.base_power_mode == MDR_BASE_INTERPOLATION) {  // shared-memory copy of the interpolation grid (see InterpGrid)
    InterpGrid* g = reinterpret_cast<InterpGrid*>(smem_raw + p.off_grid);
    if (tid < MDR_INTERP_DIMS) g->interp_dims[tid] = p.interp_dims[tid];
    for (int i = tid; i < MDR_INTERP_DIMS * MDR_INTERP_MAX_AXIS; i += blockDim.x)  // (a CTA may have fewer than 120 threads)
      (&g->interp_axes[0][0])[i] = (&p.interp_axes[0][0])[i];
  }
  if (tid == 0) {
    ctl.due_n = 0;
    for (int i = 0; i <= ring_mask; ++i) {
      mbar_init(&ctl.full[i], 1);
      mbar_init(&ctl.empty[i], p.house_warps);
    }
    mbar_init(&sctl.xbar[0], p.house_warps);  // local arrivals; the peers' contributions are transaction bytes
    mbar_init(&sctl.xbar[1], p.house_warps);
    mbar_init(&sctl.rbar, ncl * p.house_warps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // no CTA may push into a peer whose mbarriers are not initialised yet
  cluster_sync_all();
  if (warp >= p.house_warps) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int B = p.pro_batch;
    while (B > 1 && (B >> 1) * (int)gridDim.x >= p.n_tiles) B >>= 1;
    for (int it0 = 0; blockIdx.x + it0 * gridDim.x < p.n_tiles; it0 += B) prologue_pass(p, it0, B);
    return;
  }

  // ---------------- per-thread constants of the tile loop ------------------------------------
  const int N = p.N, S = p.cl_slice;
  const int C = kC > 0 ? kC : p.C;
  const int half = C >> 1;
  const int T = p.hmax;                      // house threads of the CTA (multiple of 32)
  const int nw = p.house_warps;
  const int lo = rank * S;                   // first house of this CTA's slice (index within the env)
  const int H = min(S, N - lo);              // houses of this CTA (the last rank may own fewer)
  const bool active = tid < H;
  const int ws = S + C;                      // message window: [half halo | S houses | C - half halo]
  float4* const msg0 = reinterpret_cast<float4*>(smem_raw + p.off_msg) + (half + tid);
  // halo pushes: my first C - half houses are the upper halo of the previous rank (after ITS last house), my last
  // `half` houses the lower halo of the next rank; ranks wrap around the env
  const int prev = rank == 0 ? ncl - 1 : rank - 1, next = rank == ncl - 1 ? 0 : rank + 1;
  const int H_prev = min(S, N - prev * S);
  const bool push_hi = active && tid < C - half, push_lo = active && tid >= H - half;
  const uint32_t win_base = smem_u32(smem_raw + p.off_msg);
  const uint32_t hi_addr = mapa_shared(win_base + (uint32_t)(half + H_prev + tid) * 16u, prev);
  const uint32_t lo_addr = mapa_shared(win_base + (uint32_t)(tid - (H - half)) * 16u, next);
  const uint32_t xbar_prev = mapa_shared(smem_u32(&sctl.xbar[0]), prev), xbar_next = mapa_shared(smem_u32(&sctl.xbar[0]), next);
  // bytes this CTA receives per tile: cl x nw partial sums (x 6 with metric partials) + the C halo messages
  const uint32_t tx_bytes = (uint32_t)(ncl * nw) * (kMetrics ? 24u : 4u) + (uint32_t)C * 16u;
  // warp-partial power sums of the whole env: [2][cl][nw] (+ metric partials [2][cl][nw][5])
  float* const part_base = reinterpret_cast<float*>(smem_raw + p.off_pw);
  float* const met_base = reinterpret_cast<float*>(smem_raw + p.off_met);
  const int part_buf = ncl * nw;
  const uint32_t part_saddr = smem_u32(part_base) + (uint32_t)(rank * nw + warp) * 4u;
  const uint32_t met_saddr = smem_u32(met_base) + (uint32_t)(rank * nw + warp) * 20u;
  const uint32_t xbar_saddr = smem_u32(&sctl.xbar[0]);
  float* const row = reinterpret_cast<float*>(smem_raw + p.off_stage) + tid * p.F;
  PipeEnv* const s_env = reinterpret_cast<PipeEnv*>(smem_raw + p.off_env);
  unsigned char* const in_a = smem_raw + p.off_in + tid * 16;
  unsigned char* const in_t = smem_raw + p.off_in + T * 32 + tid * 8;
  unsigned char* const in_h = smem_raw + p.off_in + T * 48 + tid * 4;
  const int in_stride = p.in_stride;
  const bool interp_mode = p.base_power_mode == MDR_BASE_INTERPOLATION;
  const int n_tiles = p.n_tiles;
  const int tile_stride = gridDim.x;
  const float inv_norm = p.f_inv_norm_reg_sig;

  // tile -> env (all CTAs of the cluster walk the same envs: gridDim.x is a multiple of cl)
  auto tile_env = [&](int tile) { return tile / ncl; };
  auto issue_tile = [&](int tile, int s) {
    if (active) {
      const unsigned h = (unsigned)tile_env(tile) * (unsigned)N + (unsigned)(lo + tid);
      const int so = s * in_stride;
      cp_async_16(in_a + so, reinterpret_cast<const float4*>(p.coef_a) + h);
      cp_async_16(in_a + so + T * 16, reinterpret_cast<const float4*>(p.coef_b) + h);
      cp_async_8(in_t + so, reinterpret_cast<const float2*>(p.temps) + h);
      cp_async_8(in_t + so + T * 8, reinterpret_cast<const float2*>(p.coef_c) + h);
      cp_async_4(in_h + so, p.hvac + h);
    }
  };
  auto fetch_action = [&](int tile) -> int {
    if (kAct == MDR_ACT_ARRAY && active) return p.actions[(unsigned)tile_env(tile) * (unsigned)N + (unsigned)(lo + tid)];
    return 0;
  };

  int tile = blockIdx.x;
  int cmd_next = 0;
  int n_refresh = 0;  // refreshes of this CTA so far (phase of rbar; identical in every CTA of the cluster)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  MDR_CTA_STAMP(1);
  if (tile < n_tiles) {
    issue_tile(tile, 0);
    cmd_next = fetch_action(tile);
  }
  cp_async_commit();

  for (int it = 0; tile < n_tiles; ++it, tile += tile_stride) {
    const int sbuf = it & 1;
    const int e = tile_env(tile);
    const unsigned h = (unsigned)e * (unsigned)N + (unsigned)(lo + tid);
    int cmd = cmd_next;
    const int next_tile = tile + tile_stride;
    if (next_tile < n_tiles) {
      issue_tile(next_tile, sbuf ^ 1);
      cmd_next = fetch_action(next_tile);
    }
    cp_async_commit();
    const int slot = it & ring_mask;
    const PipeEnv* const pe = s_env + slot;  // one env per tile
    MDR_STAMP(0);
    mbar_wait(&ctl.full[slot], (it >> ring_shift) & 1);
    const float od_old = pe->od_old;
    const float gain = pe->gain;
    MDR_STAMP(1);
    cp_async_wait<1>();
    MDR_STAMP(2);

    // ---------------- phase A: per house ---------------------------------------------------
    float t_air = 0, t_mass = 0, target = 0, p_on = 0, deadband = 0, inv_lock = 1, pen = 0, pw = 0, terr = 0;
    int on = 0, lock = 0, sso = 0;
    float4* const msg = msg0 + sbuf * ws;
    if (active) {
      const int so = sbuf * in_stride;
      const float4 ca4 = *reinterpret_cast<const float4*>(in_a + so);
      const float4 cb = *reinterpret_cast<const float4*>(in_a + so + T * 16);
      const float2 tt = *reinterpret_cast<const float2*>(in_t + so);
      const float2 cc = *reinterpret_cast<const float2*>(in_t + so + T * 8);
      const int hv = *reinterpret_cast<const int*>(in_h + so);
      target = cb.w; p_on = cb.z; deadband = cc.x;
      inv_lock = inv_real(cc.y);
      on = hv & 1; sso = hv >> 2;
      if (kAct == MDR_ACT_ARRAY) cmd = cmd != 0;
      else if (kAct == MDR_ACT_BANGBANG) cmd = tt.x > target;  // agents/bangbang_controllers.py:50-61
      else cmd = philox4x32(h + p.house_base, (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32), STREAM_ACT, p.seed).x & 1;
      // HVAC.step, :475-492
      const int dt = p.dt;
      const int lockdur = (int)cc.y;
      if (!on) sso += dt;
      lock = !(on || sso >= lockdur);
      const int new_on = lock ? 0 : cmd;
      if (!lock && new_on) sso = 0;
      if (!lock && !new_on && sso + dt < lockdur) lock = 1;
      on = new_on;
      // SingleHouse.update_temperature, :681-738, with the OLD outdoor temperature
      const float qa = (on ? cb.y : 0.0f) + gain;
      const float tss = od_old + qa * cb.x;
      const float x = tt.x - tss, y = tt.y - tss;
      t_air = tt.x + (ca4.x * x + ca4.y * y);
      t_mass = tt.y + (ca4.z * x + ca4.w * y);
      pw = on ? p_on : 0.0f;
      // SingleHouse.message :624-662 normalised as utils.py:842-868 (sso is scaled by the receiver)
      const float4 m = make_float4((t_air - target) * 0.2f, (float)sso, pw * inv_norm, p_on * inv_norm);
      msg[0] = m;
      const uint32_t boff = (uint32_t)(sbuf * ws) * 16u;
      if (push_hi) st_async_v4(hi_addr + boff, m, xbar_prev + (uint32_t)sbuf * 8u);
      if (push_lo) st_async_v4(lo_addr + boff, m, xbar_next + (uint32_t)sbuf * 8u);
      // utils.deadbandL2, utils.py:1266-1274
      const float hi = target + deadband * 0.5f, lw = target - deadband * 0.5f;
      if (hi < t_air) pen = (t_air - hi) * (t_air - hi);
      else if (lw > t_air) pen = (lw - t_air) * (lw - t_air);
      terr = t_air - target;
    }
    // warp partials -> every CTA of the cluster (own included), then one arrive per peer
    {
      const float psum = warp_sum_f(pw);
      float s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
      if (kMetrics) {
        s0 = warp_sum_f(pen); s1 = warp_sum_f(terr); s2 = warp_sum_f(fabsf(terr)); s3 = warp_sum_f(terr * terr);
        s4 = warp_max_f(fabsf(terr));
      }
      __syncwarp();  // this warp's window entries precede lane 0's (release) arrive on the CTA's own mbarrier
      if (lane == 0) {
        const uint32_t pb = (uint32_t)(sbuf * part_buf) * 4u, mb = (uint32_t)(sbuf * part_buf) * 20u;
        for (int r = 0; r < ncl; ++r) {
          const uint32_t rbar = mapa_shared(xbar_saddr + (uint32_t)sbuf * 8u, r);
          st_async_f32(mapa_shared(part_saddr + pb, r), psum, rbar);
          if (kMetrics) {
            const uint32_t ma = mapa_shared(met_saddr + mb, r);
            st_async_f32(ma, s0, rbar); st_async_f32(ma + 4, s1, rbar); st_async_f32(ma + 8, s2, rbar);
            st_async_f32(ma + 12, s3, rbar); st_async_f32(ma + 16, s4, rbar);
          }
        }
        if (warp == 0) mbar_arrive_expect_tx(&sctl.xbar[sbuf], tx_bytes);
        else mbar_arrive(&sctl.xbar[sbuf]);
      }
    }
    if (active) {
      reinterpret_cast<float2*>(p.temps)[h] = make_float2(t_air, t_mass);
      p.hvac[h] = (sso << 2) | (lock << 1) | on;
    }
    MDR_STAMP(3);
    // the staging rows of this warp may still be read by the previous tile's bulk store
    if (kObs && it > 0) {
      if (lane == 0) bulk_wait_read_all();
      __syncwarp();
    }
    MDR_STAMP(4);
    // the tile's only rendezvous: every warp of every CTA of the env has delivered its messages and partials
    mbar_wait(&sctl.xbar[sbuf], (it >> 1) & 1);
    MDR_STAMP(5);

    float P = 0;
    {
      // <= 8 x 7 partials: every lane takes its share, the shuffle tree adds them (exact: integer-valued watts < 2^24)
      const float* part = part_base + sbuf * part_buf;
      float v = 0;
      for (int i = lane; i < part_buf; i += 32) v += part[i];
      P = warp_sum_f(v);
    }
    // per-env outputs, once per env (see the header): rank 0's first thread, from the record
    const bool env_head = rank == 0 && tid == 0;
    double sig_new = pe->sig_new;
    float f_sig = pe->f_sig;
    if (interp_mode && pe->due) {  // CTA- and cluster-uniform
      sig_new = split_refresh(p, e, *pe, T, n_refresh++);
      f_sig = (float)(sig_new * p.inv_norm_sig_agents);
    }
    if (env_head) {
      p.cluster_power[e] = (double)P;
      p.t_epoch[e] = (int64_t)pe->t_new;
      p.od_temp[e] = pe->od_new;
      if (p.solar) p.solar_gain[e] = (double)pe->gain;
      p.signal[e] = sig_new;
      if (interp_mode) {
        double* s_val = reinterpret_cast<double*>(smem_raw + p.off_val);
        p.base_power[e] = pe->due ? s_val[T] : pe->base;
        p.time_since_interp[e] = pe->due ? 0 : p.time_since_interp[e] + p.dt;  // (peers read it before the rendezvous)
      } else {
        p.base_power[e] = pe->base;
      }
    }
    if (kMetrics && env_head) {
      const float* d = met_base + sbuf * part_buf * 5;
      double t[4] = {0.0, 0.0, 0.0, 0.0};
      float mxf = 0.0f;
      for (int i = 0; i < part_buf; ++i) {
        t[0] += (double)d[i * 5 + 0]; t[1] += (double)d[i * 5 + 1]; t[2] += (double)d[i * 5 + 2]; t[3] += (double)d[i * 5 + 3];
        mxf = fmaxf(mxf, d[i * 5 + 4]);
      }
      const double mx = (double)mxf, Pd = (double)P;
      const double dn = (Pd - pe->s_old) * p.inv_n;
      double* m = p.metrics + (size_t)e * MDR_N_METRICS;
      // (reductions at L2 instead of load-add-store round trips, see metrics_signal_terms)
      atomicAdd(m + MDR_M_STEPS, 1.0);
      atomicAdd(m + MDR_M_SUM_MEAN_REWARD, -(t[0] * p.inv_n * p.k_temp + dn * dn * p.k_sig));
      atomicAdd(m + MDR_M_SUM_MEAN_TEMP_OFFSET, t[1] * p.inv_n);
      atomicAdd(m + MDR_M_SUM_MEAN_TEMP_ERROR, t[2] * p.inv_n);
      atomicAdd(m + MDR_M_SUM_SQ_TEMP_ERROR, t[3]);
      atomicAdd(m + MDR_M_SUM_SQ_MAX_TEMP_ERROR, mx * mx);
      atomicMax(reinterpret_cast<unsigned long long*>(m + MDR_M_MAX_TEMP_ERROR), (unsigned long long)__double_as_longlong(mx));
      atomicAdd(m + MDR_M_SUM_OD_TEMP, pe->od_new);
      atomicAdd(m + MDR_M_SUM_CONSUMPTION, Pd);
      metrics_signal_terms(m, sig_new, Pd);
    }
    if (kObs && active) {
      row[0] = (t_air - 20.0f) * 0.2f;
      row[1] = (t_mass - 20.0f) * 0.2f;
      row[2] = (target - 20.0f) * 0.2f;
      row[3] = deadband;
      row[4] = p_on * p.f_cop_over_def_cap;
      row[5] = (float)on;
      row[6] = (float)lock;
      row[7] = (float)sso * inv_lock;
      row[8] = 1.0f;
      row[9] = f_sig;
      row[10] = (float)((double)P * p.inv_norm_sig_agents);
      const float4* win = msg - half;
      float* mrow = row + 11;
#pragma unroll
      for (int k = 0; k < (kC > 0 ? kC : C); ++k) {
        const float4 m = win[k + (k >= half ? 1 : 0)];
        mrow[4 * k + 0] = m.x;
        mrow[4 * k + 1] = m.y * inv_lock;
        mrow[4 * k + 2] = m.z;
        mrow[4 * k + 3] = m.w;
      }
    }
    if (active) {
      // reg_signal_penalty :244-247 with the OLD signal; weighting :364-372
      const float dn = (float)((double)P - pe->s_old) * p.f_inv_n;
      if (p.reward != nullptr) reinterpret_cast<float*>(p.reward)[h] = -(pen * p.f_k_temp + dn * dn * p.f_k_sig);
    }
    if (kObs) {
      const int wrow0 = warp * 32;
      const int nrows_w = min(32, H - wrow0);
      if (nrows_w > 0) {
        const int F = p.F;
        float* dst = reinterpret_cast<float*>(p.obs) + (size_t)((unsigned)e * (unsigned)N + (unsigned)(lo + wrow0)) * F;
        const float* src = reinterpret_cast<const float*>(smem_raw + p.off_stage) + wrow0 * F;
        const uint32_t bytes = (uint32_t)(nrows_w * F * sizeof(float));
        const bool bulk_ok = ((reinterpret_cast<uintptr_t>(dst) | bytes) & 15) == 0;
        if (bulk_ok) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) bulk_store_s2g(dst, src, bytes);
        } else {
          __syncwarp();
          for (int i = lane; i < nrows_w * F; i += 32) dst[i] = src[i];
          __syncwarp();
        }
      }
    }
    MDR_STAMP(6);
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctl.empty[slot]);
    MDR_STAMP(7);
    // Programmatic dependent launch is triggered when this CTA starts its LAST tile: the next step's clusters are
    // scheduled while this grid drains.  (At kernel entry -- as the single-CTA kernel does -- the early clusters of the
    // next grid take SM slots from this grid's: measured 68 vs 50 us per step on 1000 x 1000 houses.)
    if (tile + 2 * tile_stride >= n_tiles) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  cp_async_wait<0>();
  if (kObs && lane == 0) bulk_wait_read_all();
  MDR_CTA_STAMP(2);
  MDR_CTA_STAMP(3);
}


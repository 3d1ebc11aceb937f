// C ABI of the step path (include/mdr_b200.h): argument validation, launch geometry, launches,
// error mapping.  No device allocation; the only process-wide state is the launchers' mutex/atomic-protected
// per-device attribute caches and two tuning environment variables read once; no stream synchronisation
// except in the *_host entry points.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mdr_kernels.h"

using mdr::Geometry;
using mdr::KernelParams;

static thread_local char g_cuda_error[256] = "";

static int cuda_fail(cudaError_t err) {
  snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s", cudaGetErrorName(err), cudaGetErrorString(err));
  return MDR_ERR_CUDA;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int mdr_version(void) { return MDR_ABI_VERSION; }

extern "C" const char* mdr_last_cuda_error(void) { return g_cuda_error; }

extern "C" const char* mdr_strerror(int status) {
  switch (status) {
    case MDR_OK: return "ok";
    case MDR_ERR_NULL: return "a required pointer is NULL";
    case MDR_ERR_SHAPE: return "inconsistent shape (n_envs / n_houses / n_comm / n_features)";
    case MDR_ERR_MODE: return "unknown mode value";
    case MDR_ERR_ALIGN: return "buffer is not 16-byte aligned";
    case MDR_ERR_CUDA: return "CUDA runtime error (see mdr_last_cuda_error)";
    case MDR_ERR_UNSUPPORTED: return "configuration not supported by this build";
    case MDR_ERR_VERSION: return "MdrConfig.abi_version does not match the library";
    default: return "unknown status";
  }
}

extern "C" int mdr_obs_width(const MdrConfig* c) {
  if (!c) return MDR_ERR_NULL;
  // utils.normStateDict, utils.py:740-880
  int own = 11;  // T_air, T_mass, target, deadband, cap, on, lockout, sso, lockout_dur, signal, power
  if (c->state_flags & MDR_STATE_THERMAL) own += 1 + 4;
  if (c->state_flags & MDR_STATE_DAY) own += 2;
  if (c->state_flags & MDR_STATE_HOUR) own += 2;
  if (c->state_flags & MDR_STATE_SOLAR) own += 1;
  if (c->state_flags & MDR_STATE_HVAC) own += 2;
  int msg = 4;
  if (c->msg_flags & MDR_MSG_THERMAL) msg += 4;
  if (c->msg_flags & MDR_MSG_HVAC) msg += 3;
  return own + msg * c->n_comm;
}

extern "C" int mdr_validate(const MdrConfig* c) {
  if (!c) return MDR_ERR_NULL;
  if (c->abi_version != MDR_ABI_VERSION) return MDR_ERR_VERSION;
  if (c->precision != MDR_F32 && c->precision != MDR_F64) return MDR_ERR_MODE;
  if (c->n_envs < 1 || c->n_houses < 1 || c->n_comm < 0 || c->time_step < 1) return MDR_ERR_SHAPE;
  if (c->n_comm > 0 && c->comm_mode == MDR_COMM_NEIGHBOURS && c->n_comm > c->n_houses - 1) return MDR_ERR_SHAPE;
  if (c->comm_mode < MDR_COMM_NEIGHBOURS || c->comm_mode > MDR_COMM_NONE) return MDR_ERR_MODE;
  if (c->comm_mode == MDR_COMM_NONE && c->n_comm != 0) return MDR_ERR_SHAPE;
  if (c->temp_penalty_mode < MDR_PEN_INDIVIDUAL_L2 || c->temp_penalty_mode > MDR_PEN_MIXTURE) return MDR_ERR_MODE;
  if (c->base_power_mode != MDR_BASE_CONSTANT && c->base_power_mode != MDR_BASE_INTERPOLATION) return MDR_ERR_MODE;
  if (c->signal_mode < MDR_SIG_FLAT || c->signal_mode > MDR_SIG_PERLIN) return MDR_ERR_MODE;
  if (c->action_source < MDR_ACT_ARRAY || c->action_source > MDR_ACT_GREEDY) return MDR_ERR_MODE;
  if (c->n_sinusoids < 0 || c->n_sinusoids > MDR_MAX_SINUSOIDS) return MDR_ERR_SHAPE;
  if (c->n_features != mdr_obs_width(c)) return MDR_ERR_SHAPE;
  /* the on-device greedy controller sorts an env inside one CTA */
  if (c->n_houses > MDR_MAX_HOUSES_PER_ENV && c->action_source == MDR_ACT_GREEDY) return MDR_ERR_UNSUPPORTED;
  /* 32-bit house / observation indexing inside the kernel */
  if ((double)c->n_envs * c->n_houses * (c->n_features > 0 ? c->n_features : 1) >= 4294967296.0) return MDR_ERR_UNSUPPORTED;
  if (c->base_power_mode == MDR_BASE_INTERPOLATION) {
    if (c->interp_nb_agents < 1 || c->interp_nb_agents > MDR_MAX_HOUSES_PER_ENV) return MDR_ERR_SHAPE;
    for (int d = 0; d < MDR_INTERP_DIMS; ++d)
      if (c->interp_dims[d] < 2 || c->interp_dims[d] > MDR_INTERP_MAX_AXIS) return MDR_ERR_SHAPE;
    if (c->interp_dims[7] > 16) return MDR_ERR_SHAPE;
  }
  return MDR_OK;
}

static void fill_config(KernelParams& k, const MdrConfig* c);
static bool split_geometry(const MdrConfig* c, int* slice, int* k);

// envs beyond a thread-block cluster take the three-launch path through a global workspace (mdr_big.cuh)
static bool needs_big_path(const MdrConfig* c) {
  if (c->n_houses <= MDR_MAX_HOUSES_PER_ENV) return false;
  int slice = 0, k = 0;
  return c->n_houses > MDR_MAX_HOUSES_PER_CLUSTER || !split_geometry(c, &slice, &k);
}

extern "C" int mdr_workspace_bytes(const MdrConfig* cfg, size_t* bytes) {
  if (!cfg || !bytes) return MDR_ERR_NULL;
  int st = mdr_validate(cfg);
  if (st != MDR_OK) return st;
  // big path: per-env records + per-CTA totals; pipelined kernels: due-tile queue + in-order tile claiming area of a
  // launch (pipe_ws_bytes), twice (the host-buffer pipeline runs two slices concurrently)
  size_t need = needs_big_path(cfg) ? mdr::big_workspace(cfg->n_envs, cfg->n_houses) : 0;
  if (!needs_big_path(cfg)) {
    const size_t q = 2 * (size_t)mdr::pipe_ws_bytes(cfg->n_envs);
    if (q > need) need = q;
  }
  *bytes = need;
  return MDR_OK;
}

// One env split over a thread-block cluster (N > 224): CTA capacity 224 / 480 / 992 houses with a dedicated prologue
// warp, 1024 without; the smallest capacity that needs at most 8 (portable), else 16 CTAs.  The slice of rank 0 must
// hold every sampled house of an interpolation refresh.
static bool split_geometry(const MdrConfig* c, int* slice, int* k) {
  const int N = c->n_houses;
  const int need = c->base_power_mode == MDR_BASE_INTERPOLATION ? (N < c->interp_nb_agents ? N : c->interp_nb_agents) : 1;
  // A handful of envs (config 1: ONE cluster of 1000 houses) is latency-bound: the shorter the slice a CTA owns, the
  // shorter its row assembly -- up to 8 CTAs of >= 96 houses (>= the sampled houses of a refresh).
  if (c->n_envs * 8 <= 148) {
    const int floor_s = need > 96 ? need : 96;
    int kk = N / floor_s;
    if (kk > 8) kk = 8;
    if (kk >= 2) {
      int S = ((N + kk - 1) / kk + 3) & ~3;
      if (S <= 992 && N - (kk - 1) * S >= 1 && S >= need) {
        *slice = S;
        *k = kk;
        return true;
      }
    }
  }
  static const int forced_k = [] {  // experiments: MDR_SPLIT_K forces the cluster size of envs that fit (read once)
    const char* s = getenv("MDR_SPLIT_K");
    const int v = s ? atoi(s) : 0;
    return v >= 2 && v <= 16 ? v : 0;
  }();
  if (forced_k) {
    int S = ((N + forced_k - 1) / forced_k + 3) & ~3;
    if (S <= 1024 && N - (forced_k - 1) * S >= 1 && S >= need) {
      *slice = S;
      *k = forced_k;
      return true;
    }
  }
  const int caps[4] = {224, 480, 992, 1024};
  for (int max_k = 8; max_k <= 16; max_k += 8)
    for (int i = 0; i < 4; ++i) {
      const int kk = (N + caps[i] - 1) / caps[i];
      if (kk < 2 || kk > max_k) continue;
      int S = (N + kk - 1) / kk;
      S = (S + 3) & ~3;  // slices start on a multiple of 4 houses: 16-byte aligned observation rows whenever N % 4 == 0
      if (S > caps[i] || N - (kk - 1) * S < 1 || S < need) continue;
      *slice = S;
      *k = kk;
      return true;
    }
  return false;
}

// what the config alone says about the pipelined kernels' layout conditions (replayed message drops are a run-time
// input: they send a split env to the generic cluster kernel)
static bool pipe_layout_config(const MdrConfig* c) {
  return c->precision == MDR_F32 && c->comm_mode == MDR_COMM_NEIGHBOURS && c->state_flags == 0 && c->msg_flags == 0 &&
         c->temp_penalty_mode == MDR_PEN_INDIVIDUAL_L2 && c->action_source != MDR_ACT_GREEDY && !(c->flags & MDR_FLAG_NO_PIPELINE);
}

static int choose_geometry(const MdrConfig* c, bool has_obs, Geometry* g, bool need_met = false, bool allow_split = false) {
  const int N = c->n_houses, E = c->n_envs, F = c->n_features, rb = c->precision;
  static const int target_threads = [] {  // tuning knob (house threads per CTA), read once
    const char* s = getenv("MDR_TARGET_THREADS");
    const int v = s ? atoi(s) : 0;
    return v >= 32 && v <= 992 ? v : 224;  // + 32 for the prologue warp = 256
  }();
  // Splitting pays when whole-env CTAs would leave SMs idle (fewer envs than SMs) -- each CTA of this non-persistent
  // kernel is latency-bound (~7 us), so with many envs more, smaller CTAs only add waves (measured: 1000 x 1000 houses
  // 47 us as 1000 CTAs of 1024 threads, 87 us as 5000 CTAs of 256) -- and it is mandatory above 1024 houses.
  int slice = 0, ncl = 1;
  // ... unless the persistent pipelined split kernel takes the env (fp32, default layout, <= 8 x 224 houses, no drops): its
  // CTAs loop over tiles with everything prefetched, so more, smaller tiles cost nothing.
  const bool pipe_split_cfg = pipe_layout_config(c) && !(c->comm_defect_prob > 0.0) && N > 224 && N <= 8 * 224;
  const bool want_split = N > MDR_MAX_HOUSES_PER_ENV || (N > 224 && E < 148) || pipe_split_cfg;
  const bool split = allow_split && want_split && !(c->flags & MDR_FLAG_NO_CLUSTER && N <= MDR_MAX_HOUSES_PER_ENV) &&
                     c->action_source != MDR_ACT_GREEDY && split_geometry(c, &slice, &ncl);
  if (!split && N > MDR_MAX_HOUSES_PER_ENV) return MDR_ERR_UNSUPPORTED;
  int gmax = target_threads / N;
  if (gmax < 1) gmax = 1;
  if (gmax > E) gmax = E;
  int G = gmax;
  if (has_obs && !split) {  // a CTA's first observation row should be 16-byte aligned for the bulk store
    for (int k = gmax; k >= 1; --k)
      if (((size_t)k * N * F * rb) % 16 == 0) { G = k; break; }
  }
  if (split) G = 1;
  const int house_threads = ((((split ? slice : G * N) + 31) / 32)) * 32;
  const int house_warps = house_threads / 32;
  // a dedicated prologue warp when the CTA has room for one, else warp 0 runs the prologue first
  const bool extra = house_threads + 32 <= 1024;
  const int threads = house_threads + (extra ? 32 : 0);
  const int nwarps = house_warps;  // staging tiles exist for the house warps only
  const int part_stride = (N + 31) / 32 + 1;
  const int part_slots = split ? house_warps : part_stride;  // warp partials a CTA keeps per env
  const bool need_val = c->base_power_mode == MDR_BASE_INTERPOLATION || c->action_source == MDR_ACT_GREEDY;
  const bool need_pen = c->temp_penalty_mode != MDR_PEN_INDIVIDUAL_L2;
  int blocks_per_sm = 768 / threads;
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  const size_t budget = (size_t)MDR_MAX_SMEM_BYTES / blocks_per_sm - 1024;
  int rpp = 0;
  size_t smem = 0;
  for (int r = 32; r >= 1; r >>= 1) {
    smem = mdr::step_smem_layout(nullptr, rb, house_threads, G, nwarps, r, F, need_val, need_pen, has_obs, c->n_comm, part_stride, need_met, part_slots);
    if (smem <= budget) { rpp = r; break; }
  }
  if (rpp == 0) {
    for (int r = 32; r >= 1; r >>= 1) {
      smem = mdr::step_smem_layout(nullptr, rb, house_threads, G, nwarps, r, F, need_val, need_pen, has_obs, c->n_comm, part_stride, need_met, part_slots);
      if (smem <= (size_t)MDR_MAX_SMEM_BYTES) { rpp = r; break; }
    }
  }
  if (rpp == 0) return MDR_ERR_UNSUPPORTED;
  g->envs_per_cta = G;
  g->threads = threads;
  g->hmax = house_threads;
  g->house_warps = house_warps;
  g->pro_warp = extra ? house_warps : 0;
  g->part_stride = part_stride;
  g->ctas = split ? E * ncl : (E + G - 1) / G;
  g->cluster = split ? ncl : 1;
  g->cluster_slice = split ? slice : 0;
  g->part_slots = part_slots;
  g->rows_per_pass = rpp;
  g->smem_bytes = smem;
  g->pipe_smem_bytes = 0;
  g->pro_batch = 1;
  g->l2_window_base = c->l2_window_base;
  g->l2_window_bytes = c->l2_window_base ? (size_t)c->l2_window_bytes : 0;
  g->l2_hit_ratio = (float)(c->l2_hit_ratio > 0.0 && c->l2_hit_ratio <= 1.0 ? c->l2_hit_ratio : 1.0);
  g->max_ctas = c->max_ctas > 0 ? c->max_ctas : 0;
  g->no_pdl = (c->flags & MDR_FLAG_NO_PDL) != 0;
  static const bool env_static_tiles = [] { const char* s = getenv("MDR_STATIC_TILES"); return s && atoi(s) != 0; }();  // A/B runs
  g->static_tiles = (c->flags & MDR_FLAG_STATIC_TILES) != 0 || env_static_tiles;
  if (rb == MDR_F32 && extra && threads <= 256 && rpp == 32 && !split) {
    g->pro_batch = mdr::pipe_pro_batch(G, has_obs);
    const size_t ps = mdr::pipe_smem_layout(nullptr, house_threads, G, N, F, need_val, has_obs, c->n_comm, part_stride, g->pro_batch);
    if (ps <= (size_t)MDR_MAX_SMEM_BYTES) g->pipe_smem_bytes = ps;
  }
  if (rb == MDR_F32 && extra && threads <= 256 && split && ncl <= 8) {
    g->pro_batch = mdr::pipe_pro_batch(1, has_obs);
    const size_t ps = mdr::pipe_split_smem_layout(nullptr, house_threads, slice, F, need_val, has_obs, c->n_comm, ncl, house_warps,
                                                  g->pro_batch);
    if (ps <= (size_t)MDR_MAX_SMEM_BYTES) g->pipe_smem_bytes = ps;
  }
  return MDR_OK;
}

extern "C" int mdr_launch_geometry(const MdrConfig* cfg, int has_obs, int32_t* envs_per_cta, int32_t* threads,
                                   int32_t* ctas, size_t* smem_bytes, int32_t* pipelined, int32_t* cluster_size) {
  int st = mdr_validate(cfg);
  if (st != MDR_OK) return st;
  Geometry g;
  if (needs_big_path(cfg)) {
    const int nparts = (cfg->n_houses + 255) / 256;
    if (envs_per_cta) *envs_per_cta = 1;
    if (threads) *threads = 256;
    if (ctas) *ctas = cfg->n_envs * nparts;
    if (smem_bytes) *smem_bytes = 0;
    if (pipelined) *pipelined = 0;
    if (cluster_size) *cluster_size = 0;  /* 0 = plain CTAs, three launches per step */
    return MDR_OK;
  }
  st = choose_geometry(cfg, has_obs != 0, &g, false, true);
  if (st != MDR_OK) return st;
  if (cluster_size) *cluster_size = g.cluster;
  if (envs_per_cta) *envs_per_cta = g.envs_per_cta;
  if (threads) *threads = g.threads;
  if (ctas) *ctas = g.ctas;
  KernelParams k;
  fill_config(k, cfg);
  const bool pipe = !(cfg->flags & MDR_FLAG_NO_PIPELINE) &&
                    (mdr::pipe_eligible(k, g, cfg->precision) || (k.comm_defect_prob <= 0.0 && mdr::pipe_split_eligible(k, g, cfg->precision)));
  if (smem_bytes) *smem_bytes = pipe ? g.pipe_smem_bytes : g.smem_bytes;
  if (pipelined) *pipelined = pipe ? 1 : 0;
  if (!has_obs && !(cfg->flags & MDR_FLAG_NO_PIPELINE) && mdr::wide_eligible(k)) {  // one CTA per env (mdr_wide.cuh)
    if (envs_per_cta) *envs_per_cta = 1;
    if (threads) *threads = 256;
    if (ctas) *ctas = cfg->n_envs;
    if (cluster_size) *cluster_size = 1;
    if (pipelined) *pipelined = 2;
  }
  return MDR_OK;
}

static void fill_config(KernelParams& k, const MdrConfig* c) {
  memset(&k, 0, sizeof(k));
  k.E = c->n_envs; k.N = c->n_houses; k.C = c->n_comm; k.F = c->n_features; k.dt = c->time_step;
  k.comm_mode = c->comm_mode; k.state_flags = c->state_flags; k.msg_flags = c->msg_flags;
  k.temp_penalty_mode = c->temp_penalty_mode; k.solar = c->solar_gain != 0;
  k.base_power_mode = c->base_power_mode; k.signal_mode = c->signal_mode; k.n_sinusoids = c->n_sinusoids;
  k.interp_update_period = c->interp_update_period; k.interp_nb_agents = c->interp_nb_agents;
  k.perlin_nb_octaves = c->perlin_nb_octaves; k.perlin_octaves_step = c->perlin_octaves_step;
  k.action_source = c->action_source; k.seed = c->seed;
  k.alpha_temp = c->alpha_temp; k.alpha_sig = c->alpha_sig;
  k.norm_temp_penalty = c->norm_temp_penalty; k.norm_sig_penalty = c->norm_sig_penalty;
  k.mix_alpha_ind = c->mix_alpha_ind; k.mix_alpha_common = c->mix_alpha_common; k.mix_alpha_max = c->mix_alpha_max;
  k.inv_norm_reg_sig = 1.0 / c->norm_reg_sig;
  const int agents = c->obs_norm_agents > 0 ? c->obs_norm_agents : c->n_houses;
  k.inv_norm_sig_agents = 1.0 / (c->norm_reg_sig * agents);
  k.cop_over_def_cap = c->hvac_cop / c->def_cap;
  k.inv_perlin_period = c->perlin_period > 0 ? 1.0 / c->perlin_period : 0.0;
  k.inv_n = 1.0 / c->n_houses;
  k.od_amplitude = (c->day_temp - c->night_temp) / 2;
  k.od_bias = (c->day_temp + c->night_temp) / 2;
  k.two_pi_over_24 = 2 * 3.141592653589793 / 24;
  k.k_temp = c->alpha_temp / c->norm_temp_penalty;
  k.k_sig = c->alpha_sig / c->norm_sig_penalty;
  k.f_inv_norm_reg_sig = (float)k.inv_norm_reg_sig; k.f_inv_norm_sig_agents = (float)k.inv_norm_sig_agents;
  k.f_cop_over_def_cap = (float)k.cop_over_def_cap; k.f_inv_n = (float)k.inv_n;
  k.f_k_temp = (float)k.k_temp; k.f_k_sig = (float)k.k_sig;
  k.def_ua = c->def_ua; k.def_cm = c->def_cm; k.def_ca = c->def_ca; k.def_hm = c->def_hm;
  k.def_cop = c->def_cop; k.def_latent = c->def_latent; k.def_cap = c->def_cap;
  k.hvac_cop = c->hvac_cop; k.hvac_latent = c->hvac_latent;
  k.day_temp = c->day_temp; k.night_temp = c->night_temp; k.temp_std = c->temp_std;
  k.window_area = c->window_area; k.shading_coeff = c->shading_coeff;
  k.avg_power_per_hvac = c->avg_power_per_hvac;
  memcpy(k.sin_periods, c->sin_periods, sizeof(k.sin_periods));
  memcpy(k.sin_ratios, c->sin_ratios, sizeof(k.sin_ratios));
  k.steps_amplitude_per_hvac = c->steps_amplitude_per_hvac; k.steps_period = c->steps_period;
  k.perlin_amplitude = c->perlin_amplitude; k.perlin_period = c->perlin_period;
  k.comm_defect_prob = c->comm_defect_prob;
  memcpy(k.interp_dims, c->interp_dims, sizeof(k.interp_dims));
  memcpy(k.interp_axes, c->interp_axes, sizeof(k.interp_axes));
}

static int fill_houses(KernelParams& k, const MdrConfig* c, const MdrHouses* h, bool for_step) {
  if (!h) return MDR_ERR_NULL;
  if (!h->coef_a || !h->coef_b || !h->coef_c) return MDR_ERR_NULL;
  if (!aligned16(h->coef_a) || !aligned16(h->coef_b) || !aligned16(h->coef_c)) return MDR_ERR_ALIGN;
  const bool interp = c->base_power_mode == MDR_BASE_INTERPOLATION;
  if (interp && !h->interp_key) return MDR_ERR_NULL;
  if (for_step) {
    if (!h->temps || !h->hvac) return MDR_ERR_NULL;
    if (!aligned16(h->temps) || !aligned16(h->hvac)) return MDR_ERR_ALIGN;
    const bool need_raw = (c->state_flags & MDR_STATE_THERMAL) || (c->msg_flags & MDR_MSG_THERMAL);
    if (need_raw && (!h->ua || !h->cm || !h->ca || !h->hm)) return MDR_ERR_NULL;
    if ((c->msg_flags & MDR_MSG_HVAC) && !h->cap) return MDR_ERR_NULL;
    if (c->action_source == MDR_ACT_GREEDY && !h->cap) return MDR_ERR_NULL;
  } else {
    if (!h->ua || !h->cm || !h->ca || !h->hm || !h->cap || !h->target || !h->deadband || !h->lockout_dur)
      return MDR_ERR_NULL;
  }
  k.ua = h->ua; k.cm = h->cm; k.ca = h->ca; k.hm = h->hm; k.cap = h->cap;
  k.target = h->target; k.deadband = h->deadband; k.lockout_dur = h->lockout_dur;
  k.coef_a = h->coef_a; k.coef_b = h->coef_b; k.coef_c = h->coef_c; k.interp_key = h->interp_key;
  k.temps = h->temps; k.hvac = h->hvac;
  return MDR_OK;
}

static int fill_step(KernelParams& k, const MdrConfig* c, const MdrEnvs* e, const MdrStepInputs* in,
                     const MdrOutputs* out, int is_reset) {
  if (!e || !in || !out) return MDR_ERR_NULL;
  if (!e->t_epoch || !e->phase || !e->od_temp || !e->artificial_ratio || !e->max_power || !e->base_power ||
      !e->signal || !e->cluster_power)
    return MDR_ERR_NULL;
  if (c->solar_gain && !e->solar_gain) return MDR_ERR_NULL;
  const bool interp = c->base_power_mode == MDR_BASE_INTERPOLATION;
  if (interp && (!e->time_since_interp || !in->interp_table)) return MDR_ERR_NULL;
  if (c->signal_mode == MDR_SIG_PERLIN && !in->signal_noise && !e->perlin_seed) return MDR_ERR_NULL;
  if (!is_reset && c->action_source == MDR_ACT_ARRAY && !in->actions) return MDR_ERR_NULL;
  if ((c->comm_mode == MDR_COMM_TABLE || c->comm_mode == MDR_COMM_TABLE_PER_ENV) && c->n_comm > 0 && !in->comm_table &&
      out->obs)
    return MDR_ERR_NULL;
  if (out->obs && !aligned16(out->obs)) return MDR_ERR_ALIGN;
  k.metrics = e->metrics;
  k.workspace = e->workspace;
  k.t_epoch = e->t_epoch; k.phase = e->phase; k.od_temp = e->od_temp; k.solar_gain = e->solar_gain;
  k.artificial_ratio = e->artificial_ratio; k.max_power = e->max_power;
  k.base_power = e->base_power; k.signal = e->signal; k.cluster_power = e->cluster_power;
  k.time_since_interp = e->time_since_interp; k.perlin_seed = e->perlin_seed;
  k.actions = in->actions; k.od_noise = in->od_noise; k.signal_noise = in->signal_noise;
  k.interp_ids = in->interp_ids; k.msg_keep = in->msg_keep; k.comm_table = in->comm_table;
  if (in->env_mask) {
    if (is_reset != 1 || out->obs) return MDR_ERR_UNSUPPORTED;  // masked reset only, observation via mdr_observe
    k.env_mask = in->env_mask;
  }
  k.interp_table = in->interp_table; k.step_index = in->step_index; k.step_counter = in->step_counter;
  k.obs = out->obs; k.reward = out->reward;
  k.is_reset = is_reset;
  return MDR_OK;
}

extern "C" int mdr_l2_persist_limit(int device, size_t bytes, size_t* granted_bytes, size_t* max_window_bytes) {
  cudaError_t err = cudaSetDevice(device);
  if (err != cudaSuccess) return cuda_fail(err);
  int max_persist = 0, max_window = 0;
  err = cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
  if (err != cudaSuccess) return cuda_fail(err);
  err = cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
  if (err != cudaSuccess) return cuda_fail(err);
  const size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
  err = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
  if (err != cudaSuccess) return cuda_fail(err);
  if (bytes == 0) {
    err = cudaCtxResetPersistingL2Cache();
    if (err != cudaSuccess) return cuda_fail(err);
  }
  size_t got = 0;
  err = cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
  if (err != cudaSuccess) return cuda_fail(err);
  if (granted_bytes) *granted_bytes = got;
  if (max_window_bytes) *max_window_bytes = (size_t)max_window;
  return MDR_OK;
}

extern "C" int mdr_populate(const MdrConfig* cfg, const MdrPopulationSpec* spec, const MdrHouses* h, const MdrEnvs* e,
                            const uint8_t* env_mask, uint64_t draw_index, void* stream) {
  int st = mdr_validate(cfg);
  if (st != MDR_OK) return st;
  if (!spec || !h || !e) return MDR_ERR_NULL;
  if (!h->ua || !h->cm || !h->ca || !h->hm || !h->cap || !h->target || !h->deadband || !h->lockout_dur || !h->temps || !h->hvac)
    return MDR_ERR_NULL;
  if (!e->t_epoch || !e->phase || !e->od_temp || !e->artificial_ratio || !e->max_power || !e->base_power || !e->signal ||
      !e->cluster_power)
    return MDR_ERR_NULL;
  if (spec->n_cap < 0 || spec->n_cap > 8 || spec->lockout_noise < 0 || spec->lockout_duration - spec->lockout_noise < 0)
    return MDR_ERR_SHAPE;  // "Lockout duration must be positive", env/MA_DemandResponse.py:438-461
  KernelParams k;
  fill_config(k, cfg);
  k.temps = h->temps; k.hvac = h->hvac;
  k.t_epoch = e->t_epoch; k.phase = e->phase; k.od_temp = e->od_temp; k.solar_gain = e->solar_gain;
  k.artificial_ratio = e->artificial_ratio; k.max_power = e->max_power; k.base_power = e->base_power;
  k.signal = e->signal; k.cluster_power = e->cluster_power; k.time_since_interp = e->time_since_interp;
  k.perlin_seed = e->perlin_seed;
  cudaError_t err = cudaSetDevice(cfg->device);
  if (err != cudaSuccess) return cuda_fail(err);
  err = mdr::launch_populate(k, *spec, env_mask, const_cast<double*>(h->ua), const_cast<double*>(h->cm),
                             const_cast<double*>(h->ca), const_cast<double*>(h->hm), const_cast<double*>(h->cap),
                             const_cast<double*>(h->target), const_cast<double*>(h->deadband),
                             const_cast<int32_t*>(h->lockout_dur), cfg->precision, draw_index, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? MDR_OK : cuda_fail(err);
}

extern "C" int mdr_precompute(const MdrConfig* cfg, const MdrHouses* houses, void* stream) {
  int st = mdr_validate(cfg);
  if (st != MDR_OK) return st;
  KernelParams k;
  fill_config(k, cfg);
  st = fill_houses(k, cfg, houses, false);
  if (st != MDR_OK) return st;
  cudaError_t err = cudaSetDevice(cfg->device);
  if (err != cudaSuccess) return cuda_fail(err);
  err = mdr::launch_precompute_any(k, cfg->precision, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? MDR_OK : cuda_fail(err);
}

static int run_steps(const MdrConfig* cfg, const MdrHouses* houses, const MdrEnvs* envs, const MdrStepInputs* in,
                     const MdrOutputs* out, int n_steps, int is_reset, cudaStream_t stream, int env_base = 0, int ws_envs = 0) {
  int st = mdr_validate(cfg);
  if (st != MDR_OK) return st;
  if (n_steps < 1) return MDR_ERR_SHAPE;
  KernelParams k;
  fill_config(k, cfg);
  k.env_base = env_base;
  k.house_base = (unsigned)env_base * (unsigned)cfg->n_houses;
  st = fill_houses(k, cfg, houses, true);
  if (st != MDR_OK) return st;
  st = fill_step(k, cfg, envs, in, out, is_reset);
  if (st != MDR_OK) return st;
  if (needs_big_path(cfg)) {
    if (!k.workspace) return MDR_ERR_NULL;          // mdr_workspace_bytes() of device memory in MdrEnvs.workspace
    if (k.metrics) return MDR_ERR_UNSUPPORTED;      // accumulate from the returned tensors instead
    if (!aligned16(k.workspace)) return MDR_ERR_ALIGN;
    cudaError_t err = cudaSetDevice(cfg->device);
    if (err != cudaSuccess) return cuda_fail(err);
    for (int i = 0; i < n_steps; ++i) {
      err = mdr::launch_big(k, cfg->precision, k.workspace, stream);
      if (err != cudaSuccess) return cuda_fail(err);
      k.step_index += 1;
    }
    return MDR_OK;
  }
  // Geometry: whole envs per CTA (N <= 1024) is what the fused multi-step kernel needs; otherwise an env of more
  // than 224 houses is split over a thread-block cluster -- or, without an observation, walked by one CTA (mdr_wide.cuh).
  Geometry g;
  bool fused = false;
  if (cfg->n_houses <= MDR_MAX_HOUSES_PER_ENV) {
    st = choose_geometry(cfg, out->obs != nullptr, &g, k.metrics != nullptr, false);
    if (st != MDR_OK) return st;
    fused = !(cfg->flags & MDR_FLAG_NO_FUSED) && mdr::fused_eligible(k) && g.pro_warp >= g.house_warps &&
            (n_steps > 1 || k.metrics != nullptr);
  }
  if (!fused && !(cfg->flags & MDR_FLAG_NO_PIPELINE) && mdr::wide_eligible(k)) {
    cudaError_t werr = cudaSetDevice(cfg->device);
    if (werr != cudaSuccess) return cuda_fail(werr);
    for (int i = 0; i < n_steps; ++i) {
      werr = mdr::launch_wide(k, cfg->precision, (cfg->flags & MDR_FLAG_NO_PDL) != 0, stream);
      if (werr != cudaSuccess) return cuda_fail(werr);
      k.step_index += 1;
    }
    return MDR_OK;
  }
  if (!fused) {
    st = choose_geometry(cfg, out->obs != nullptr, &g, k.metrics != nullptr, true);
    if (st != MDR_OK) return st;
  }
  k.G = g.envs_per_cta;
  k.hmax = g.hmax;
  k.rows_per_pass = g.rows_per_pass;
  k.div_magic = (unsigned)(4294967296ull / (unsigned)cfg->n_houses) + 1u;
  {
    int L = 16;  // nb_octaves + 1 items per env in production mode
    while (L > 1 && L * g.envs_per_cta > 32) L >>= 1;
    k.pro_lanes = L;
  }
  k.house_warps = g.house_warps;
  k.house_threads = g.house_warps * 32;
  k.ns = cfg->n_houses + cfg->n_comm;
  k.in_stride = g.hmax * 52;
  k.pro_warp = g.pro_warp;
  k.part_stride = g.part_stride;
  k.cl = g.cluster;
  k.cl_slice = g.cluster_slice;
  mdr::step_smem_layout(&k, cfg->precision, g.hmax, g.envs_per_cta, g.house_warps, g.rows_per_pass, cfg->n_features,
                        cfg->base_power_mode == MDR_BASE_INTERPOLATION || cfg->action_source == MDR_ACT_GREEDY,
                        cfg->temp_penalty_mode != MDR_PEN_INDIVIDUAL_L2, out->obs != nullptr, cfg->n_comm, g.part_stride,
                        k.metrics != nullptr, g.part_slots);
  cudaError_t err = cudaSetDevice(cfg->device);
  if (err != cudaSuccess) return cuda_fail(err);
  const bool pipe_split = !(cfg->flags & MDR_FLAG_NO_PIPELINE) && mdr::pipe_split_eligible(k, g, cfg->precision);
  if (pipe_split) {
    k.pro_batch = g.pro_batch;
    k.in_stride = k.hmax * 52;
    mdr::pipe_split_smem_layout(&k, k.hmax, g.cluster_slice, cfg->n_features, cfg->base_power_mode == MDR_BASE_INTERPOLATION,
                                out->obs != nullptr, cfg->n_comm, g.cluster, g.house_warps, g.pro_batch);
  }
  const bool pipe = !(cfg->flags & MDR_FLAG_NO_PIPELINE) && mdr::pipe_eligible(k, g, cfg->precision);
  if (pipe) {
    // the due queue needs its full size; a caller that passes less scratch (or none) gets the per-CTA refresh
    if (k.workspace && !aligned16(k.workspace)) return MDR_ERR_ALIGN;
    k.pro_batch = g.pro_batch;
    // in-order tile claiming needs the scratch area; without it (or on request) every CTA walks a fixed tile list
    // (the layout follows the n_envs the workspace was sized for: slices of the host pipeline share it with whole steps)
    const int cap = ws_envs > 0 ? ws_envs : cfg->n_envs;
    k.dyn_off = (k.workspace && !g.static_tiles) ? (int)mdr::due_queue_bytes(cap) : 0;
    k.dyn_list_off = k.dyn_off + 64 + (int)mdr::dyn_flags_bytes(cap);
    k.dyn_rec_off = k.dyn_list_off + (int)mdr::dyn_flags_bytes(cap);
    int L = 16;  // lanes per env within one tile's lane group
    while (L > 1 && L * g.envs_per_cta * g.pro_batch > 32) L >>= 1;
    k.pro_lanes = L;
    k.in_stride = k.hmax * 52;
    mdr::pipe_smem_layout(&k, k.hmax, g.envs_per_cta, cfg->n_houses, cfg->n_features,
                          cfg->base_power_mode == MDR_BASE_INTERPOLATION, out->obs != nullptr, cfg->n_comm, g.part_stride,
                          g.pro_batch);
  }
  if (fused) {
    err = mdr::launch_fused(k, g, cfg->precision, n_steps, stream);
    return err == cudaSuccess ? MDR_OK : cuda_fail(err);
  }
  for (int i = 0; i < n_steps; ++i) {
    err = pipe_split ? mdr::launch_pipe_split(k, g, stream)
          : pipe     ? mdr::launch_pipe(k, g, stream)
                     : mdr::launch_step_any(k, g, cfg->precision, stream);
    if (err != cudaSuccess) return cuda_fail(err);
    k.step_index += 1;
  }
  return MDR_OK;
}

extern "C" int mdr_sample_actions(const float* probs, int64_t n_rows, int32_t n_actions, uint64_t seed, uint64_t draw_index,
                                  const uint64_t* draw_counter, uint8_t* actions, float* chosen_prob, void* stream) {
  if (!probs || !actions) return MDR_ERR_NULL;
  if (n_rows < 0 || n_actions < 1 || n_actions > 256) return MDR_ERR_SHAPE;
  if (n_rows == 0) return MDR_OK;
  cudaError_t err = mdr::launch_sample_actions(probs, (long long)n_rows, n_actions, seed, draw_index, draw_counter, actions,
                                               chosen_prob, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? MDR_OK : cuda_fail(err);
}

extern "C" int mdr_reset(const MdrConfig* cfg, const MdrHouses* houses, const MdrEnvs* envs, const MdrStepInputs* in,
                         const MdrOutputs* out, void* stream) {
  return run_steps(cfg, houses, envs, in, out, 1, 1, static_cast<cudaStream_t>(stream));
}

extern "C" int mdr_observe(const MdrConfig* cfg, const MdrHouses* houses, const MdrEnvs* envs, const MdrStepInputs* in,
                           const MdrOutputs* out, void* stream) {
  return run_steps(cfg, houses, envs, in, out, 1, 2, static_cast<cudaStream_t>(stream));
}

extern "C" int mdr_step(const MdrConfig* cfg, const MdrHouses* houses, const MdrEnvs* envs, const MdrStepInputs* in,
                        const MdrOutputs* out, int32_t n_steps, void* stream) {
  return run_steps(cfg, houses, envs, in, out, n_steps, 0, static_cast<cudaStream_t>(stream));
}

namespace mdr {
// entry points of the host-buffer pipeline (mdr_host.cu): one step / the compact observation record of a slice of the
// env axis (`env_base` keeps the Philox keys those of the whole shard)
int run_steps_slice(const MdrConfig* cfg, const MdrHouses* h, const MdrEnvs* e, const MdrStepInputs* in, const MdrOutputs* out,
                    int env_base, int ws_envs, cudaStream_t stream) {
  return run_steps(cfg, h, e, in, out, 1, 0, stream, env_base, ws_envs);
}
int compact_slice(const MdrConfig* cfg, const MdrHouses* h, const MdrEnvs* e, void* out, cudaStream_t stream) {
  KernelParams k;
  fill_config(k, cfg);
  k.temps = h->temps; k.hvac = h->hvac; k.coef_b = h->coef_b; k.coef_c = h->coef_c;
  k.signal = e->signal; k.cluster_power = e->cluster_power;
  cudaError_t err = launch_compact_obs(k, cfg->precision, out, stream);
  return err == cudaSuccess ? MDR_OK : cuda_fail(err);
}
}  // namespace mdr

extern "C" int mdr_step_host(const MdrConfig* cfg, const MdrHouses* houses, const MdrEnvs* envs,
                             const MdrStepInputs* in, const MdrOutputs* out, const uint8_t* host_actions,
                             void* host_obs, void* host_reward, double* host_power, double* host_signal,
                             MdrHostCtx* ctx, void* stream_v) {
  if (!cfg || !in || !out || !envs || !houses) return MDR_ERR_NULL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  cudaError_t err = cudaSetDevice(cfg->device);
  if (err != cudaSuccess) return cuda_fail(err);
  const size_t houses_total = (size_t)cfg->n_envs * cfg->n_houses;
  // pipelined path: slices of the env axis over two streams, compact observation records expanded on the host
  if (ctx != nullptr && host_obs != nullptr && mdr::host_compact_eligible(cfg, in, out) && !needs_big_path(cfg)) {
    int st = mdr_validate(cfg);
    if (st != MDR_OK) return st;
    return mdr::host_pipeline_step(ctx, cfg, houses, envs, in, out, host_actions, host_obs, host_reward, host_power,
                                   host_signal, stream, cuda_fail);
  }
  if (host_actions) {
    if (!in->actions) return MDR_ERR_NULL;
    err = cudaMemcpyAsync(const_cast<uint8_t*>(in->actions), host_actions, houses_total, cudaMemcpyHostToDevice, stream);
    if (err != cudaSuccess) return cuda_fail(err);
  }
  int st = run_steps(cfg, houses, envs, in, out, 1, 0, stream);
  if (st != MDR_OK) return st;
  const size_t rb = (size_t)cfg->precision;
  if (host_obs && out->obs) {
    err = cudaMemcpyAsync(host_obs, out->obs, houses_total * cfg->n_features * rb, cudaMemcpyDeviceToHost, stream);
    if (err != cudaSuccess) return cuda_fail(err);
  }
  if (host_reward && out->reward) {
    err = cudaMemcpyAsync(host_reward, out->reward, houses_total * rb, cudaMemcpyDeviceToHost, stream);
    if (err != cudaSuccess) return cuda_fail(err);
  }
  if (host_power) {
    err = cudaMemcpyAsync(host_power, envs->cluster_power, sizeof(double) * cfg->n_envs, cudaMemcpyDeviceToHost, stream);
    if (err != cudaSuccess) return cuda_fail(err);
  }
  if (host_signal) {
    err = cudaMemcpyAsync(host_signal, envs->signal, sizeof(double) * cfg->n_envs, cudaMemcpyDeviceToHost, stream);
    if (err != cudaSuccess) return cuda_fail(err);
  }
  err = cudaStreamSynchronize(stream);
  return err == cudaSuccess ? MDR_OK : cuda_fail(err);
}

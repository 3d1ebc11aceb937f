/*
 * mdr_b200.h -- C ABI of the B200-native MADemandResponseEnv step path.
 *
 * The reference (zhimaerfan/marl-demandresponse-original) is pure Python and has no
 * FFI/plugin interface for this path: its boundary is the duck-typed class
 * `MADemandResponseEnv` (env/MA_DemandResponse.py:37) with `reset()` (:135-172) and
 * `step(action_dict)` (:174-210).  This header is what a ctypes binding of that class binds
 * instead of the Python object graph HVAC -> SingleHouse -> ClusterHouses -> PowerGrid; each
 * entry point cites the reference code it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every device buffer is allocated by the caller (torch,
 *     cudaMalloc, ...) on `MdrConfig.device`; the library never allocates or frees device memory (except the staging
 *     buffers of an explicitly created MdrHostCtx) and never synchronises the stream except in mdr_step_host.  The only process-wide state is a
 *     mutex-protected cache of per-device launch attributes (occupancy, opt-in shared memory) and the
 *     tuning environment variables MDR_TARGET_THREADS / MDR_PRO_BATCH, read once at first use;
 *   - every function returns an MdrStatus (0 = OK, negative = error) and never throws/prints;
 *   - "real" is float when MdrConfig.precision == MDR_F32 and double when == MDR_F64;
 *   - per-house arrays are [n_envs * n_houses], env-major (house h of env e at e*n_houses+h),
 *     16-byte aligned; per-env arrays are [n_envs];
 *   - callable from any host thread; re-entrant; one stream per shard.
 */
#ifndef MDR_B200_H
#define MDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDR_ABI_VERSION 10
#define MDR_MAX_SINUSOIDS 8
#define MDR_INTERP_DIMS 10
#define MDR_INTERP_MAX_AXIS 12
#define MDR_MAX_HOUSES_PER_ENV 1024      /* largest env that lives in ONE CTA (thread per house) */
#define MDR_MAX_HOUSES_PER_CLUSTER 16384 /* largest env split over one thread-block cluster (16 CTAs x 1024 houses) */

typedef enum MdrStatus {
  MDR_OK = 0,
  MDR_ERR_NULL = -1,        /* a required pointer is NULL */
  MDR_ERR_SHAPE = -2,       /* n_envs / n_houses / n_comm / n_features inconsistent */
  MDR_ERR_MODE = -3,        /* unknown enum value (reference raises ValueError) */
  MDR_ERR_ALIGN = -4,       /* a buffer is not 16-byte aligned */
  MDR_ERR_CUDA = -5,        /* a CUDA runtime call failed (see mdr_last_cuda_error) */
  MDR_ERR_UNSUPPORTED = -6, /* valid in the reference but outside this build (e.g. greedy controller with N > 1024) */
  MDR_ERR_VERSION = -7      /* MdrConfig.abi_version != MDR_ABI_VERSION */
} MdrStatus;

enum { MDR_F32 = 4, MDR_F64 = 8 };

/* agents_comm_mode, env/MA_DemandResponse.py:806-902 */
enum {
  MDR_COMM_NEIGHBOURS = 0, /* implicit circular table, :816-828 */
  MDR_COMM_TABLE = 1,      /* explicit int32 [n_houses, n_comm] shared by all envs
                              (closed_groups :830-844, random_fixed :849-854, neighbours_2D :856-890) */
  MDR_COMM_TABLE_PER_ENV = 2, /* explicit int32 [n_envs, n_houses, n_comm] (random_sample :976-983, replayed) */
  MDR_COMM_NONE = 3        /* no_message :893-895 */
};

/* state_properties / message_properties, utils.py:740-880 */
enum {
  MDR_STATE_HOUR = 1, MDR_STATE_DAY = 2, MDR_STATE_SOLAR = 4, MDR_STATE_THERMAL = 8, MDR_STATE_HVAC = 16,
  MDR_MSG_THERMAL = 1, MDR_MSG_HVAC = 2
};

/* temp_penalty_mode, env/MA_DemandResponse.py:253-328 */
enum { MDR_PEN_INDIVIDUAL_L2 = 0, MDR_PEN_COMMON_L2 = 1, MDR_PEN_COMMON_MAX = 2, MDR_PEN_MIXTURE = 3 };
/* base_power_mode, env/MA_DemandResponse.py:1248-1255 */
enum { MDR_BASE_CONSTANT = 0, MDR_BASE_INTERPOLATION = 1 };
/* signal_mode, env/MA_DemandResponse.py:1257-1310 */
enum { MDR_SIG_FLAT = 0, MDR_SIG_SINUSOIDALS = 1, MDR_SIG_REGULAR_STEPS = 2, MDR_SIG_PERLIN = 3 };
/* where the HVAC commands come from */
enum {
  MDR_ACT_ARRAY = 0,    /* MdrStepInputs.actions (the reference's action_dict) */
  MDR_ACT_BANGBANG = 1, /* on-device restatement of agents/bangbang_controllers.py:41-61 (benchmark source) */
  MDR_ACT_RANDOM = 2,   /* Philox Bernoulli(1/2) per house (benchmark source) */
  MDR_ACT_GREEDY = 3    /* on-device restatement of agents/greedy_myopic_controller.py:29-49 (needs MdrHouses.cap) */
};

/* Flattened form of the reference's nested config dict (config.py), built once per env. */
typedef struct MdrConfig {
  int32_t abi_version; /* MDR_ABI_VERSION */
  int32_t device;      /* CUDA device ordinal that owns every buffer */
  int32_t precision;   /* MDR_F32 | MDR_F64 */
  int32_t n_envs;      /* E: independent clusters in this shard */
  int32_t n_houses;    /* N: default_env_prop.cluster_prop.nb_agents */
  int32_t n_comm;      /* C: messages per agent (min(nb_agents_comm, N-1), :808-810) */
  int32_t n_features;  /* F: must equal mdr_obs_width(cfg) */
  int32_t time_step;   /* seconds, default_env_prop.time_step */
  int32_t comm_mode;   /* MDR_COMM_* */
  int32_t state_flags; /* MDR_STATE_* bitmask */
  int32_t msg_flags;   /* MDR_MSG_* bitmask */
  int32_t temp_penalty_mode; /* MDR_PEN_* */
  int32_t solar_gain;  /* default_house_prop.solar_gain_bool */
  int32_t base_power_mode;   /* MDR_BASE_* */
  int32_t signal_mode; /* MDR_SIG_* */
  int32_t n_sinusoids;
  int32_t interp_update_period; /* seconds */
  int32_t interp_nb_agents;     /* houses sampled per refresh */
  int32_t perlin_nb_octaves, perlin_octaves_step;
  int32_t action_source; /* MDR_ACT_* */
  int32_t obs_norm_agents; /* nb_agents used by normStateDict (utils.py:833-839); normally == n_houses */
  /* reward, env/MA_DemandResponse.py:330-373 */
  double alpha_temp, alpha_sig, norm_temp_penalty, norm_sig_penalty;
  double mix_alpha_ind, mix_alpha_common, mix_alpha_max;
  /* normStateDict divisors, utils.py:740-880 */
  double norm_reg_sig, def_ua, def_cm, def_ca, def_hm, def_cop, def_latent, def_cap;
  /* HVAC constants shared by all houses (the reference never noises COP / latent fraction) */
  double hvac_cop, hvac_latent;
  /* outdoor temperature model, :1057-1081 */
  double day_temp, night_temp, temp_std;
  /* solar gain, utils.py:1277-1350 */
  double window_area, shading_coeff;
  /* power grid */
  double avg_power_per_hvac;
  double sin_periods[MDR_MAX_SINUSOIDS], sin_ratios[MDR_MAX_SINUSOIDS];
  double steps_amplitude_per_hvac, steps_period;
  double perlin_amplitude, perlin_period;
  double comm_defect_prob; /* only used when MdrStepInputs.msg_keep is NULL */
  /* interpolation grid, monteCarlo/interp_parameters_dict.json in interp_dict_keys.csv order */
  int32_t interp_dims[MDR_INTERP_DIMS];
  double interp_axes[MDR_INTERP_DIMS][MDR_INTERP_MAX_AXIS];
  uint64_t seed; /* Philox key for draws that are not replayed */
  /* Optional L2 residency window for the step launches (0 bytes = none): the address range that holds the
     per-house state and coefficients re-read by every step.  Accesses inside it are launched with
     cudaAccessPropertyPersisting (hit ratio l2_hit_ratio), so the once-written observation stream does not
     evict them.  Needs a persisting-L2 carve-out on the device: see mdr_l2_persist_limit. */
  const void *l2_window_base;
  uint64_t l2_window_bytes;
  double l2_hit_ratio;
  /* Launch options (0 = defaults): MDR_FLAG_* bitmask, and a cap on the CTAs of the persistent pipelined kernel
     (0 = SMs x resident CTAs; tests use a small cap so that every CTA walks many tiles). */
  int32_t flags;
  int32_t max_ctas;
} MdrConfig;

/* MdrConfig.flags */
enum {
  MDR_FLAG_NO_PIPELINE = 1, /* never take the persistent pipelined kernel (generic kernel instead) */
  MDR_FLAG_NO_FUSED = 2,    /* never take the fused multi-step kernel (one launch per step instead) */
  MDR_FLAG_NO_PDL = 4,      /* launch without programmatic dependent launch */
  MDR_FLAG_NO_CLUSTER = 8,  /* never split an env of 225..1024 houses over a thread-block cluster (one CTA per env) */
  MDR_FLAG_STATIC_TILES = 16 /* pipelined kernel: every CTA walks a fixed, strided list of tiles (no in-order claiming) */
};

/* Per-house struct-of-arrays.  Packed vectors keep every access a coalesced 8/16-byte load. */
typedef struct MdrHouses {
  /* raw properties (double regardless of precision); read by mdr_precompute, by the optional
     observation blocks and by the interpolation refresh */
  const double *ua, *cm, *ca, *hm;    /* SingleHouse.Ua/Cm/Ca/Hm, :581-584 */
  const double *cap;                  /* HVAC.cooling_capacity, :428 */
  const double *target, *deadband;    /* SingleHouse.target_temp/deadband, :577-578 */
  const int32_t *lockout_dur;         /* HVAC.lockout_duration (noise already applied), :430 */
  /* derived coefficients written by mdr_precompute, real4/real4/real2 per house */
  void *coef_a; /* (d11, m12, m21, d22): expm(A*dt) - I of the ETP ODE, :704-735 */
  void *coef_b; /* (1/Ua, Q_on = -cap/(1+latent), P_on = cap/COP, target) */
  void *coef_c; /* (deadband, lockout_duration as real) */
  int32_t *interp_key; /* flat table offset of the nearest (Ua,Cm,Ca,Hm ratio, HVAC power) cell */
  /* state, read and written by every step */
  void *temps;   /* real2 (T_air, T_mass) in Celsius, :536-537 */
  int32_t *hvac; /* (seconds_since_off << 2) | (lockout << 1) | turned_on, :405-406,433 */
} MdrHouses;

/* Per-env (cluster + power grid) scalars; always double / integer. */
typedef struct MdrEnvs {
  int64_t *t_epoch;          /* naive seconds since 1970-01-01 of MADemandResponseEnv.datetime */
  const double *phase;       /* ClusterHouses.phase, :789-792 */
  double *od_temp;           /* ClusterHouses.current_OD_temp, :793,1037 */
  double *solar_gain;        /* SingleHouse.current_solar_gain used by the last update, :694 */
  const double *artificial_ratio; /* PowerGrid.artificial_ratio, :1116 */
  const double *max_power;   /* ClusterHouses.max_power, :796-802 */
  double *base_power;        /* PowerGrid.base_power */
  double *signal;            /* PowerGrid.current_signal */
  double *cluster_power;     /* ClusterHouses.cluster_hvac_power */
  int32_t *time_since_interp;/* PowerGrid.time_since_last_interp */
  const double *perlin_seed; /* seed of the device perlin (production mode only) */
  double *metrics;           /* optional [n_envs, MDR_N_METRICS] running accumulators (see MDR_M_*), or NULL */
  void *workspace;           /* mdr_workspace_bytes() of device scratch, 16-byte aligned (0 bytes needed -> may be NULL) */
} MdrEnvs;

/* Per-env running accumulators of MdrEnvs.metrics ([n_envs, MDR_N_METRICS] doubles, += by every mdr_step call
   that takes the fused multi-step path).  They are the quantities main-deploy.py:124-209 and
   metrics.py:22-47 accumulate, before their final divisions; T = new air temperature, S = new signal,
   P = cluster power of the step, r = reward, N = houses per env. */
enum {
  MDR_M_STEPS = 0,                /* steps accumulated */
  MDR_M_SUM_MEAN_REWARD = 1,      /* sum_t sum_k r_k / N              (metrics.py:25) */
  MDR_M_SUM_MEAN_TEMP_OFFSET = 2, /* sum_t sum_k (T_k - target_k) / N (main-deploy.py:128) */
  MDR_M_SUM_MEAN_TEMP_ERROR = 3,  /* sum_t sum_k |T_k - target_k| / N (:129) */
  MDR_M_SUM_SQ_TEMP_ERROR = 4,    /* sum_t sum_k (T_k - target_k)^2   (:136) */
  MDR_M_SUM_SQ_MAX_TEMP_ERROR = 5,/* sum_t (max_k |T_k - target_k|)^2 (:139) */
  MDR_M_MAX_TEMP_ERROR = 6,       /* max_t max_k |T_k - target_k|     (:130-131), a running max, not a sum */
  MDR_M_SUM_OD_TEMP = 7,          /* sum_t OD_temp                    (:140) */
  MDR_M_SUM_SIGNAL = 8,           /* sum_t S                          (:141) */
  MDR_M_SUM_CONSUMPTION = 9,      /* sum_t P                          (:142) */
  MDR_M_SUM_SIGNAL_OFFSET = 10,   /* sum_t (S - P)                    (:144-145) */
  MDR_M_SUM_SIGNAL_ERROR = 11,    /* sum_t |S - P|                    (:146) */
  MDR_M_SUM_SQ_SIGNAL_ERROR = 12, /* sum_t (S - P)^2                  (:149) */
  MDR_N_METRICS = 13
};

/* Inputs of one step.  NULL replay pointers select on-device generation. */
typedef struct MdrStepInputs {
  const uint8_t *actions;     /* [E*N] nonzero = ON (ignored unless action_source == MDR_ACT_ARRAY) */
  const double *od_noise;     /* [E] replayed random.gauss(0, temp_std) draw (:1079) */
  const double *signal_noise; /* [E] replayed Perlin.calculate_noise value (:1299) */
  const int32_t *interp_ids;  /* [E*interp_nb_agents] replayed random.choices ids (:1214) */
  const uint8_t *msg_keep;    /* [E*N*C] 1 = delivered (replays np.random.rand() > p, :992) */
  const int32_t *comm_table;  /* MDR_COMM_TABLE(_PER_ENV) neighbour ids */
  const void *interp_table;   /* real[prod(interp_dims)], C order (mergedGridSearchResultFinal.npy) */
  uint64_t step_index;        /* counter for the Philox streams */
  const uint64_t *step_counter; /* optional DEVICE counter added to step_index by the kernels: a CUDA graph that captured
                                   a rollout advances it on the device, so replays do not repeat their draws */
  const uint8_t *env_mask;    /* mdr_reset only: [E], nonzero = reset this env; NULL = all.  Needs out->obs == NULL
                                 (follow with mdr_observe, which has no side effects) */
} MdrStepInputs;

/* Reset-time randomness of one population draw (mdr_populate): the parameters utils.applyPropertyNoise
   (utils.py:573-709) reads from default_house_prop / noise_house_prop[noise_mode] / default_hvac_prop /
   noise_hvac_prop[noise_mode] / default_env_prop, flattened. */
typedef struct MdrPopulationSpec {
  double init_air_temp, init_mass_temp, target_temp, deadband; /* default_house_prop, config.py:12-26 */
  double ua, cm, ca, hm;
  double std_start_temp, std_target_temp, factor_thermo_low, factor_thermo_high; /* noise_house_prop */
  double cap_list[8];      /* noise_hvac_prop cooling_capacity_list[default capacity] (utils.py:669-676) */
  int32_t n_cap;
  int32_t lockout_duration, lockout_noise; /* default_hvac_prop, env/MA_DemandResponse.py:430 */
  int32_t random_start;    /* start_datetime_mode == "random": + randrange(364) days + randrange(86400) s */
  int64_t start_epoch;     /* naive seconds since 1970-01-01 of default_env_prop.start_datetime */
  int32_t random_phase;    /* temp_parameters[temp_mode].random_phase_offset */
  int32_t interp_update_period;
  double artificial_ratio, artificial_ratio_range; /* power_grid_prop, :1116 */
} MdrPopulationSpec;

typedef struct MdrOutputs {
  void *obs;    /* real [E, N, F]: the normStateDict vector of every agent, or NULL to skip */
  void *reward; /* real [E, N], or NULL to skip */
} MdrOutputs;

int mdr_version(void);
const char *mdr_strerror(int status);
/* text of the last CUDA error seen by the calling thread ("" if none) */
const char *mdr_last_cuda_error(void);

/* F of utils.normStateDict (utils.py:740-880) for these flags; negative MdrStatus on error. */
int mdr_obs_width(const MdrConfig *cfg);

/* Device scratch a step of this configuration uses in MdrEnvs.workspace (zero it ONCE after allocating; the kernels
   keep it consistent from launch to launch, and it must not be shared by launches that can run concurrently):
   - envs larger than a thread-block cluster can hold (n_houses > MDR_MAX_HOUSES_PER_CLUSTER): the per-env records
     and per-CTA totals of the three-launch path (required there);
   - the persistent pipelined kernel (optional; NULL = every CTA walks a fixed list of tiles and refreshes its own due
     tiles): the hand-over records of the per-env prologue, a ready flag per tile and the counter from which the CTAs
     claim their tiles in address order, and -- with interpolated base power -- the queue through which the CTAs
     share the tiles whose table refresh is due.  Sized for two launches in flight (mdr_step_host's two streams).
     In this mode the CTAs of one launch depend on each other (a grid barrier behind the per-env prologue) and must
     be co-resident: the grid never exceeds SMs x resident CTAs, launches of one stream follow each other, and the
     hardware dispatches the CTAs of an older grid before those of a younger one.  Do not run it next to kernels that
     hold SMs indefinitely, and give envs that are stepped concurrently from streams of DIFFERENT priority
     MDR_FLAG_STATIC_TILES (two half-resident grids would wait for each other; a wait of 2 s traps instead of hanging). */
int mdr_workspace_bytes(const MdrConfig *cfg, size_t *bytes);

/* Validates cfg (modes, shapes) the way the reference constructors raise ValueError
   (env/MA_DemandResponse.py:249,324,898,1169,1306). */
int mdr_validate(const MdrConfig *cfg);

/* Launch geometry the step kernel will use (for tests and the roofline report).  `ctas` counts the
   G-env tiles; `pipelined` is 1 when a plain production-mode step of this configuration runs the
   persistent software-pipelined kernel (grid = SMs x resident CTAs, looping over the tiles), 2 when -- without an
   observation, on clusters of 225 .. 8 192 houses -- one CTA walks a whole env (no inter-CTA traffic at all);
   `cluster_size` > 1 when one env is split over the CTAs of a thread-block cluster (N > 224: cluster power and
   penalties cross the CTAs through distributed shared memory, ClusterHouses.step :1005-1055 for any nb_agents). */
int mdr_launch_geometry(const MdrConfig *cfg, int has_obs, int32_t *envs_per_cta, int32_t *threads,
                        int32_t *ctas, size_t *smem_bytes, int32_t *pipelined, int32_t *cluster_size);

/* Replaces the per-step recomputation of a,b,c,r1,r2,A3,A4,exp(r*dt) in
   SingleHouse.update_temperature (:704-735) and HVAC.get_Q/power_consumption (:494-523):
   fills coef_a/b/c and interp_key from the raw properties. */
int mdr_precompute(const MdrConfig *cfg, const MdrHouses *houses, void *stream);

/* MADemandResponseEnv.build_environment tail + reset (:133, :163-172): PowerGrid.step at the
   start datetime (initial signal, incl. the first interpolation), cluster power = 0, solar
   gain schedule, and the initial observation. */
int mdr_reset(const MdrConfig *cfg, const MdrHouses *houses, const MdrEnvs *envs,
              const MdrStepInputs *in, const MdrOutputs *out, void *stream);

/* Observation of the current state without advancing anything (what reset() returns at
   :163-172 once the signal is known; also used after loading a checkpoint). */
int mdr_observe(const MdrConfig *cfg, const MdrHouses *houses, const MdrEnvs *envs,
                const MdrStepInputs *in, const MdrOutputs *out, void *stream);

/* MADemandResponseEnv.step (:174-210) for every env of the shard, `n_steps` times.  n_steps > 1 is meant for
   on-device action/noise sources; with MDR_ACT_ARRAY or replayed noise arrays the SAME arrays are applied at every
   one of the n_steps steps.  One kernel launch per step -- except that steps which need
   nothing from the host between them (on-device action source, no replayed noise, individual_L2 penalty,
   out->obs == NULL, constant base power -- or interpolated base power with n_houses <= interp_nb_agents, no
   solar gain and a signal mode other than regular_steps) run as ONE fused launch with the house state in registers
   (the main-deploy.py:102-209 loop).  envs->metrics, when given, is accumulated by that path and by single-step
   launches of every other configuration (which then run the generic kernel, not the pipelined one). */
int mdr_step(const MdrConfig *cfg, const MdrHouses *houses, const MdrEnvs *envs,
             const MdrStepInputs *in, const MdrOutputs *out, int32_t n_steps, void *stream);

/* Device-side replacement of the host population builders: draws the raw house properties
   (houses->ua..lockout_dur), the initial state (temps, hvac) and the per-env scalars (t_epoch, phase, od_temp,
   artificial_ratio, max_power, perlin_seed; signal/base/cluster power zeroed) of every env selected by `env_mask`
   ([E] bytes, NULL = all) from Philox streams keyed by (cfg->seed, draw_index, house or env).  Distribution-level
   parity with utils.applyPropertyNoise / HVAC.__init__ / ClusterHouses.__init__ / PowerGrid.__init__; envs not
   selected are not touched.  Follow with mdr_precompute and mdr_reset (same mask). */
int mdr_populate(const MdrConfig *cfg, const MdrPopulationSpec *spec, const MdrHouses *houses, const MdrEnvs *envs,
                 const uint8_t *env_mask, uint64_t draw_index, void *stream);

/* Rollout collection: the per-agent, per-step `Categorical(action_prob).sample()` of the reference's
   learners (agents/ppo.py:68-75) for a whole batch in one launch.  probs = float [n_rows, n_actions] (un-normalised is
   fine: Categorical divides by the row sum), actions = uint8 [n_rows] (what mdr_step reads), chosen_prob = float [n_rows]
   or NULL: probs[row, action], the `a_log_prob` entry of agents/ppo.py:92-107.  Draws are Philox uniforms keyed by
   (seed, row, draw_index + *draw_counter); draw_counter is an optional device counter (see MdrStepInputs.step_counter). */
int mdr_sample_actions(const float *probs, int64_t n_rows, int32_t n_actions, uint64_t seed, uint64_t draw_index,
                       const uint64_t *draw_counter, uint8_t *actions, float *chosen_prob, void *stream);

/* Sets the device's persisting-L2 carve-out (cudaLimitPersistingL2CacheSize) to min(bytes, device maximum);
   bytes = 0 resets it and drops persisting lines.  Reports the granted carve-out and the largest access
   policy window of the device.  Device-global setting: call once per process and device. */
int mdr_l2_persist_limit(int device, size_t bytes, size_t *granted_bytes, size_t *max_window_bytes);

/* Resources of the pipelined host-buffer path (two internal streams, events, a device + a pinned host staging buffer
   of n_envs * n_houses * 16 reals, `n_threads` worker threads; 0 = the calling thread's CPU affinity count, at most 24;
   n_slices 0 = 8).  The ONE place where the library allocates; the pinned staging is first touched by the creating
   thread, so pin the process to the GPU's NUMA node before creating it.  One context per shard; not thread-safe. */
typedef struct MdrHostCtx MdrHostCtx;
int mdr_host_ctx_create(const MdrConfig *cfg, int32_t n_threads, int32_t n_slices, MdrHostCtx **ctx);
int mdr_host_ctx_destroy(MdrHostCtx *ctx);
int mdr_host_ctx_info(const MdrHostCtx *ctx, int32_t *n_threads, int32_t *n_slices, size_t *compact_bytes);

/* Same step with HOST buffers: copies `host_actions` to `in->actions` (device staging), runs the step, brings
   obs / reward / per-env (power, signal) back into the host pointers and synchronises the stream.  This is the call
   the dict API and the e2e benchmark use.
   ctx == NULL: one H2D, one launch, four D2H on `stream` (the full [E, N, F] observation crosses PCIe).
   ctx != NULL and the default observation layout (implicit `neighbours` messages, no optional feature blocks, no
   message drops): the env axis is cut into slices alternating between the context's two streams, so uploads, kernels
   and downloads of different slices overlap; per house only a 16-real record crosses PCIe (own 11 features, its
   4-real message, 1 / lockout_duration -- utils.py:842-868: 4*C of the F features of a row are copies of other
   houses' messages) and the context's threads expand the rows into `host_obs`.  The host buffers are bit-identical to
   the ctx == NULL path; `out->obs` (device) is NOT written on this path.  Other layouts fall back to the serial path. */
int mdr_step_host(const MdrConfig *cfg, const MdrHouses *houses, const MdrEnvs *envs,
                  const MdrStepInputs *in, const MdrOutputs *out, const uint8_t *host_actions,
                  void *host_obs, void *host_reward, double *host_power, double *host_signal,
                  MdrHostCtx *ctx, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MDR_B200_H */

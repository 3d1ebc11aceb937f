// Internal interface between the C ABI (mdr_abi.cu) and the kernels (mdr_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mdr_b200.h"

#define MDR_MAX_SMEM_BYTES (227 * 1024)

namespace mdr {

// Everything the kernels need, passed by value as the kernel parameter (constant bank).
struct KernelParams {
  // shape / geometry
  int E, N, C, F, G, hmax, rows_per_pass, dt;
  int off_msg, off_pw, off_val, off_pen, off_env, off_stage, off_in, off_ctl, off_met;  // shared-memory carve-up (bytes)
  int ns, house_threads, in_stride;                            // N + C, house_warps * 32, bytes of one cp.async input stage
  int pro_batch;                                              // pipelined kernel: tiles the prologue warp produces per pass
  int n_tiles;                                                // pipelined kernel: number of G-env tiles
  int n_fused;                                                // fused multi-step kernel: steps of this launch
  int precision;                                              // MDR_F32 / MDR_F64 (kernels that are not templated on it)
  int pro_lanes;                                              // lanes of the prologue warp cooperating on one env (power of two)
  int pro_warp, house_warps, part_stride;                     // prologue warp id, warps that own houses, partial-sum stride
  int off_grid;                                               // pipelined kernels: shared-memory copy of the interpolation grid
  int cl, cl_slice, off_cl;                                   // env split over a cluster of `cl` CTAs, `cl_slice` houses each; ClusterTot offset
  int dyn_off, dyn_list_off, dyn_rec_off;                     // pipelined kernel: byte offsets of the tile-claim header (0 = static tiles), the due list and the records in `workspace`
  unsigned div_magic;                                         // floor(2^32 / N) + 1: tid / N == umulhi(tid, magic)
  int is_reset, comm_mode, state_flags, msg_flags, temp_penalty_mode, solar, base_power_mode, signal_mode;
  int n_sinusoids, interp_update_period, interp_nb_agents, perlin_nb_octaves, perlin_octaves_step, action_source;
  // per-house arrays
  const double *ua, *cm, *ca, *hm, *cap, *target, *deadband;
  const int32_t* lockout_dur;
  void *coef_a, *coef_b, *coef_c;
  int32_t* interp_key;
  void* temps;
  int32_t* hvac;
  // per-env arrays
  int64_t* t_epoch;
  const double* phase;
  double *od_temp, *solar_gain;
  const double *artificial_ratio, *max_power;
  double *base_power, *signal, *cluster_power;
  int32_t* time_since_interp;
  const double* perlin_seed;
  // step inputs / outputs
  const uint8_t* actions;
  const double *od_noise, *signal_noise;
  const int32_t* interp_ids;
  const uint8_t* msg_keep;
  const uint8_t* env_mask;  // mdr_reset of a subset of the envs (generic kernel only), or nullptr
  const int32_t* comm_table;
  const void* interp_table;
  void *obs, *reward;
  double* metrics;  // [E, MDR_N_METRICS] running accumulators (fused multi-step kernel), or nullptr
  void* workspace;  // mdr_workspace_bytes() of scratch (envs beyond a thread-block cluster), or nullptr
  uint64_t step_index, seed;
  int env_base;          // index of this call's first env in the shard (host-buffer pipeline steps slices of the env axis)
  unsigned house_base;   // env_base * N
  const uint64_t* step_counter;  // optional device-resident addend of step_index (CUDA-graph replay)
  // scalars
  double alpha_temp, alpha_sig, norm_temp_penalty, norm_sig_penalty, mix_alpha_ind, mix_alpha_common, mix_alpha_max;
  double inv_norm_reg_sig, inv_norm_sig_agents, cop_over_def_cap, inv_perlin_period, inv_n, k_temp, k_sig, od_amplitude, od_bias, two_pi_over_24;
  float f_inv_norm_reg_sig, f_inv_norm_sig_agents, f_cop_over_def_cap, f_inv_n, f_k_temp, f_k_sig;  // fp32 copies (pipelined kernel)
  double def_ua, def_cm, def_ca, def_hm, def_cop, def_latent, def_cap, hvac_cop, hvac_latent;
  double day_temp, night_temp, temp_std, window_area, shading_coeff, avg_power_per_hvac;
  double sin_periods[MDR_MAX_SINUSOIDS], sin_ratios[MDR_MAX_SINUSOIDS];
  double steps_amplitude_per_hvac, steps_period, perlin_amplitude, perlin_period, comm_defect_prob;
  int interp_dims[MDR_INTERP_DIMS];
  double interp_axes[MDR_INTERP_DIMS][MDR_INTERP_MAX_AXIS];
};

struct Geometry {
  int envs_per_cta, threads, ctas, rows_per_pass, hmax, house_warps, pro_warp, part_stride, part_slots, pro_batch;
  size_t smem_bytes, pipe_smem_bytes;
  const void* l2_window_base;  // optional persisting-L2 access policy window of the launch
  size_t l2_window_bytes;
  float l2_hit_ratio;
  int cluster, cluster_slice;  // one env split over `cluster` CTAs (thread-block cluster) of `cluster_slice` houses; 1 = whole envs per CTA
  int max_ctas;  // cap on the persistent pipelined grid (0 = SMs x resident CTAs)
  bool no_pdl;   // MDR_FLAG_NO_PDL
  bool static_tiles;  // MDR_FLAG_STATIC_TILES
};

size_t step_smem_layout(KernelParams* kp, int real_bytes, int hmax, int genvs, int nwarps, int rows_per_pass,
                        int n_features, bool need_val, bool need_pen, bool has_obs, int n_comm, int part_stride, bool need_met,
                        int part_slots = 0);
size_t pipe_smem_layout(KernelParams* kp, int hmax, int genvs, int n_houses, int n_features, bool need_val, bool has_obs,
                        int n_comm, int part_stride, int pro_batch);
int pipe_pro_batch(int envs_per_cta, bool has_obs);
bool pipe_eligible(const KernelParams& kp, const Geometry& g, int precision);
cudaError_t launch_pipe(const KernelParams& kp, const Geometry& g, cudaStream_t stream);
cudaError_t launch_pipe_split(const KernelParams& kp, const Geometry& g, cudaStream_t stream);
bool pipe_split_eligible(const KernelParams& kp, const Geometry& g, int precision);
size_t pipe_split_smem_layout(KernelParams* kp, int hmax, int slice, int n_features, bool need_val, bool has_obs, int n_comm,
                              int cluster, int house_warps, int pro_batch);
cudaError_t launch_populate(const KernelParams& kp, const MdrPopulationSpec& spec, const uint8_t* env_mask, double* ua, double* cm,
                            double* ca, double* hm, double* cap, double* target, double* deadband, int32_t* lockout_dur,
                            int precision, uint64_t draw_index, cudaStream_t stream);
bool fused_eligible(const KernelParams& kp);
cudaError_t launch_big(const KernelParams& kp, int precision, void* workspace, cudaStream_t stream);
bool wide_eligible(const KernelParams& kp);  // mdr_wide.cuh: one CTA per env of 225..8192 houses, no observation
cudaError_t launch_wide(const KernelParams& kp, int precision, bool no_pdl, cudaStream_t stream);
size_t big_workspace(int n_envs, int n_houses);
inline size_t due_queue_bytes(int n_envs) { return (64 + 4 * (size_t)n_envs + 63) & ~(size_t)63; }  // header + a word per tile
// Scratch of ONE pipelined launch in MdrEnvs.workspace: [due-tile queue (strided lists) | claim header (64 B) | a due
// word per tile | the due list | a 64-byte hand-over record per env] (mdr_pipe.cuh: DueQueue, DynHdr).  Two launches
// may be in flight (host pipeline).
inline size_t dyn_flags_bytes(int n_envs) { return (4 * (size_t)n_envs + 63) & ~(size_t)63; }
inline size_t pipe_ws_bytes(int n_envs) { return due_queue_bytes(n_envs) + 64 + 2 * dyn_flags_bytes(n_envs) + 64 * (size_t)n_envs; }
cudaError_t launch_compact_obs(const KernelParams& kp, int precision, void* out, cudaStream_t stream);
cudaError_t launch_sample_actions(const float* probs, long long n_rows, int n_actions, uint64_t seed, uint64_t draw_index,
                                  const uint64_t* draw_counter, uint8_t* actions, float* chosen_prob, cudaStream_t stream);
cudaError_t launch_fused(const KernelParams& kp, const Geometry& g, int precision, int n_steps, cudaStream_t stream);
cudaError_t launch_precompute_any(const KernelParams& kp, int precision, cudaStream_t stream);
cudaError_t launch_step_any(const KernelParams& kp, const Geometry& g, int precision, cudaStream_t stream);

// host-buffer pipeline (mdr_host.cu <-> mdr_abi.cu)
bool host_compact_eligible(const MdrConfig* c, const MdrStepInputs* in, const MdrOutputs* out);
int host_pipeline_step(MdrHostCtx* ctx, const MdrConfig* cfg, const MdrHouses* h, const MdrEnvs* e, const MdrStepInputs* in,
                       const MdrOutputs* out, const uint8_t* host_actions, void* host_obs, void* host_reward,
                       double* host_power, double* host_signal, cudaStream_t user_stream, int (*fail)(cudaError_t));
int run_steps_slice(const MdrConfig* cfg, const MdrHouses* h, const MdrEnvs* e, const MdrStepInputs* in, const MdrOutputs* out,
                    int env_base, int ws_envs, cudaStream_t stream);  // ws_envs: n_envs the workspace was sized (and laid out) for
int compact_slice(const MdrConfig* cfg, const MdrHouses* h, const MdrEnvs* e, void* out, cudaStream_t stream);

}  // namespace mdr

// Launcher, eligibility and shared-memory layout of mdr::step_pipe_split_kernel (mdr_pipe_split.cuh).
// Included by mdr_kernels.cu inside namespace mdr, after the common launch helpers.
#pragma once

template <int kC, int kAct, bool kObs, bool kMetrics>
static cudaError_t launch_pipe_split_t(const KernelParams& kp_in, const Geometry& g, cudaStream_t stream) {
  static std::atomic<uint64_t> latch{0};
  static std::mutex mu;
  static std::vector<OccEntry> cache;  // ctas_per_sm holds the co-resident CLUSTERS of the device here
  auto kernel = step_pipe_split_kernel<kC, kAct, kObs, kMetrics>;
  cudaError_t err = ensure_max_smem(kernel, latch);
  if (err != cudaSuccess) return err;
  int dev = 0;
  err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = (unsigned)g.cluster;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  int n_attrs = 1;
  cudaLaunchConfig_t lc = {};
  lc.blockDim = dim3((unsigned)g.threads);
  lc.dynamicSmemBytes = g.pipe_smem_bytes;
  lc.stream = stream;
  lc.attrs = attrs;
  lc.numAttrs = 1;
  int max_clusters = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    for (const OccEntry& o : cache)
      if (o.dev == dev && o.threads == g.threads * 64 + g.cluster && o.smem == g.pipe_smem_bytes) max_clusters = o.ctas_per_sm;
    if (max_clusters == 0) {
      lc.gridDim = dim3((unsigned)g.cluster * 64u);
      err = cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &lc);
      if (err != cudaSuccess) return err;
      if (max_clusters < 1) return cudaErrorLaunchOutOfResources;
      cache.push_back(OccEntry{dev, g.threads * 64 + g.cluster, g.pipe_smem_bytes, max_clusters, 0});
    }
  }
  KernelParams kp = kp_in;
  kp.n_tiles = g.ctas;  // E * cl
  int clusters = max_clusters;
  if (g.max_ctas > 0 && clusters * g.cluster > g.max_ctas) clusters = g.max_ctas / g.cluster > 0 ? g.max_ctas / g.cluster : 1;
  if (clusters > kp.E) clusters = kp.E;
  lc.gridDim = dim3((unsigned)(clusters * g.cluster));
  static const bool verbose = getenv("MDR_VERBOSE") != nullptr;
  if (verbose) {
    static int printed = 0;
    if (printed++ < 4)
      fprintf(stderr, "[mdr] step_pipe_split_kernel: cluster %d x %d threads, smem %zu B, co-resident clusters %d -> grid %d CTAs, %d tiles\n",
              g.cluster, g.threads, g.pipe_smem_bytes, max_clusters, clusters * g.cluster, kp.n_tiles);
  }
  if (!g.no_pdl) {
    attrs[n_attrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n_attrs].val.programmaticStreamSerializationAllowed = 1;
    ++n_attrs;
  }
  lc.numAttrs = n_attrs;
  return cudaLaunchKernelEx(&lc, kernel, kp);
}

template <int kC, bool kObs, bool kMetrics>
static cudaError_t launch_pipe_split_c(const KernelParams& kp, const Geometry& g, cudaStream_t stream) {
  if (kp.action_source == MDR_ACT_ARRAY) return launch_pipe_split_t<kC, MDR_ACT_ARRAY, kObs, kMetrics>(kp, g, stream);
  if (kp.action_source == MDR_ACT_BANGBANG) return launch_pipe_split_t<kC, MDR_ACT_BANGBANG, kObs, kMetrics>(kp, g, stream);
  return launch_pipe_split_t<kC, MDR_ACT_RANDOM, kObs, kMetrics>(kp, g, stream);
}

cudaError_t launch_pipe_split(const KernelParams& kp, const Geometry& g, cudaStream_t stream) {
  const bool m = kp.metrics != nullptr;
  if (kp.obs != nullptr) {
    if (kp.C == 10) return m ? launch_pipe_split_c<10, true, true>(kp, g, stream) : launch_pipe_split_c<10, true, false>(kp, g, stream);
    return m ? launch_pipe_split_c<0, true, true>(kp, g, stream) : launch_pipe_split_c<0, true, false>(kp, g, stream);
  }
  if (kp.C == 10) return m ? launch_pipe_split_c<10, false, true>(kp, g, stream) : launch_pipe_split_c<10, false, false>(kp, g, stream);
  return m ? launch_pipe_split_c<0, false, true>(kp, g, stream) : launch_pipe_split_c<0, false, false>(kp, g, stream);
}

// same conditions as the single-CTA pipelined kernel, for an env split over <= 8 CTAs of <= 224 houses; no message drops
bool pipe_split_eligible(const KernelParams& kp, const Geometry& g, int precision) {
  return precision == MDR_F32 && kp.is_reset == 0 && kp.comm_mode == MDR_COMM_NEIGHBOURS && kp.state_flags == 0 &&
         kp.msg_flags == 0 && kp.temp_penalty_mode == MDR_PEN_INDIVIDUAL_L2 && kp.msg_keep == nullptr &&
         !(kp.comm_defect_prob > 0.0) && g.cluster >= 2 && g.cluster <= kMaxSplit && g.threads <= 256 &&
         g.pro_warp >= g.house_warps && g.pipe_smem_bytes > 0 && kp.action_source != MDR_ACT_GREEDY;
}

// shared-memory carve-up of the split kernel (one env slice of `slice` houses per tile)
size_t pipe_split_smem_layout(KernelParams* kp, int hmax, int slice, int n_features, bool need_val, bool has_obs, int n_comm,
                              int cluster, int house_warps, int pro_batch) {
  size_t o = 0;
  const size_t off_msg = o;   o += align16((size_t)2 * (slice + n_comm) * 4 * sizeof(float));
  const size_t off_pw = o;    o += align16((size_t)2 * cluster * house_warps * sizeof(float));
  const size_t off_met = o;   o += align16((size_t)2 * cluster * house_warps * 5 * sizeof(float));
  const size_t off_val = o;   o += need_val ? align16(((size_t)hmax + 2) * sizeof(double)) : 0;
  const size_t off_grid = o;  o += need_val ? align16(sizeof(InterpGrid)) : 0;
  const size_t off_env = o;   o += align16((size_t)2 * pro_batch * sizeof(PipeEnv));
  const size_t off_ctl = o;   o += align16(sizeof(SplitCtl));
  const size_t off_stage = o; o += has_obs ? align16((size_t)slice * n_features * sizeof(float)) : 0;
  const size_t off_in = o;    o += align16((size_t)2 * hmax * 52);
  if (kp) {
    kp->off_msg = (int)off_msg; kp->off_pw = (int)off_pw; kp->off_val = (int)off_val; kp->off_pen = 0;
    kp->off_env = (int)off_env; kp->off_stage = (int)off_stage; kp->off_in = (int)off_in; kp->off_ctl = (int)off_ctl;
    kp->off_met = (int)off_met; kp->off_grid = (int)off_grid;
  }
  return o;
}

// Host-buffer pipeline of mdr_step_host (include/mdr_b200.h): what a caller with HOST arrays pays is PCIe, so
//   * the env axis is cut into slices that alternate between two internal streams: while slice k's results travel to
//     the host, slice k+1's actions go up and its kernels run;
//   * with the default observation layout only a compact 16-real record per house crosses PCIe (mdr_compact.cuh) and a
//     small pool of host threads expands it into the caller's [E, N, F] buffer (non-temporal stores), slice by slice,
//     while later slices are still in flight;
//   * everything the pipeline owns (streams, events, pinned staging, worker threads) lives in an explicit MdrHostCtx.
// Bit-identical to the serial path (one H2D, one launch, four D2H): same kernels, same Philox keys (env_base), and the
// expansion multiplies the same floats the row assembly multiplies.
#include <emmintrin.h>
#include <sched.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "mdr_expand.h"
#include "mdr_kernels.h"

namespace {

constexpr int kMaxSlices = 16;

struct Pool {
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv, done_cv;
  std::function<void(int)> job;  // job(worker index)
  int generation = 0, pending = 0;
  bool stop = false;

  explicit Pool(int n) {
    for (int i = 0; i < n; ++i)
      threads.emplace_back([this, i] {
        int seen = 0;
        for (;;) {
          std::function<void(int)> j;
          {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            j = job;
          }
          j(i);
          {
            std::lock_guard<std::mutex> lk(mu);
            if (--pending == 0) done_cv.notify_all();
          }
        }
      });
  }
  void start(std::function<void(int)> j) {  // every worker runs j(index) once
    std::lock_guard<std::mutex> lk(mu);
    job = std::move(j);
    pending = (int)threads.size();
    ++generation;
    cv.notify_all();
  }
  void finish() {  // returns when all workers are done with the job of the last start()
    std::unique_lock<std::mutex> lk(mu);
    done_cv.wait(lk, [&] { return pending == 0; });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    for (auto& t : threads) t.join();
  }
};

}  // namespace

struct MdrHostCtx {
  int device = 0, n_threads = 1, n_slices = 1;
  size_t compact_bytes = 0;  // per shard: E * N * 16 * sizeof(real)
  cudaStream_t streams[2] = {nullptr, nullptr};
  cudaEvent_t ev_start = nullptr, ev_slice[kMaxSlices] = {}, ev_end[2] = {nullptr, nullptr};
  void* d_compact = nullptr;
  void* h_compact = nullptr;  // pinned
  Pool* pool = nullptr;
};

namespace mdr {

static inline const void* off(const void* p, size_t bytes) { return p ? static_cast<const char*>(p) + bytes : nullptr; }
static inline void* off(void* p, size_t bytes) { return p ? static_cast<char*>(p) + bytes : nullptr; }

bool host_compact_eligible(const MdrConfig* c, const MdrStepInputs* in, const MdrOutputs* out) {
  return c->comm_mode == MDR_COMM_NEIGHBOURS && c->state_flags == 0 && c->msg_flags == 0 && in->msg_keep == nullptr &&
         !(c->comm_defect_prob > 0.0) && c->n_houses <= MDR_MAX_HOUSES_PER_CLUSTER && in->env_mask == nullptr &&
         c->n_features == 11 + 4 * c->n_comm && out != nullptr;
}

// one step over HOST buffers through the pipeline; returns an MdrStatus
int host_pipeline_step(MdrHostCtx* ctx, const MdrConfig* cfg, const MdrHouses* h, const MdrEnvs* e, const MdrStepInputs* in,
                       const MdrOutputs* out, const uint8_t* host_actions, void* host_obs, void* host_reward,
                       double* host_power, double* host_signal, cudaStream_t user_stream, int (*fail)(cudaError_t)) {
  const int E = cfg->n_envs, N = cfg->n_houses, C = cfg->n_comm;
  const size_t rb = (size_t)cfg->precision;
  if ((size_t)E * N * 16 * rb > ctx->compact_bytes) return MDR_ERR_SHAPE;
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) return fail(err_); } while (0)
  CK(cudaSetDevice(cfg->device));
  // slices start on multiples of 16 envs: every per-house array stays 16-byte aligned at the slice boundary
  int n_slices = ctx->n_slices;
  int per = ((E + n_slices - 1) / n_slices + 15) & ~15;
  if (per < 16) per = 16;
  n_slices = (E + per - 1) / per;
  CK(cudaEventRecord(ctx->ev_start, user_stream));
  CK(cudaStreamWaitEvent(ctx->streams[0], ctx->ev_start, 0));
  CK(cudaStreamWaitEvent(ctx->streams[1], ctx->ev_start, 0));
  for (int s = 0; s < n_slices; ++s) {
    const int e0 = s * per, ne = (E - e0 < per ? E - e0 : per);
    const size_t h0 = (size_t)e0 * N;
    cudaStream_t st = ctx->streams[s & 1];
    MdrConfig c = *cfg;
    c.n_envs = ne;
    c.l2_window_base = nullptr;
    c.l2_window_bytes = 0;
    MdrHouses hs = *h;
    hs.ua = (const double*)off(h->ua, h0 * 8); hs.cm = (const double*)off(h->cm, h0 * 8);
    hs.ca = (const double*)off(h->ca, h0 * 8); hs.hm = (const double*)off(h->hm, h0 * 8);
    hs.cap = (const double*)off(h->cap, h0 * 8); hs.target = (const double*)off(h->target, h0 * 8);
    hs.deadband = (const double*)off(h->deadband, h0 * 8); hs.lockout_dur = (const int32_t*)off(h->lockout_dur, h0 * 4);
    hs.coef_a = off(h->coef_a, h0 * 4 * rb); hs.coef_b = off(h->coef_b, h0 * 4 * rb); hs.coef_c = off(h->coef_c, h0 * 2 * rb);
    hs.interp_key = (int32_t*)off(h->interp_key, h0 * 4); hs.temps = off(h->temps, h0 * 2 * rb);
    hs.hvac = (int32_t*)off(h->hvac, h0 * 4);
    MdrEnvs es = *e;
    es.t_epoch = (int64_t*)off(e->t_epoch, (size_t)e0 * 8); es.phase = (const double*)off(e->phase, (size_t)e0 * 8);
    es.od_temp = (double*)off(e->od_temp, (size_t)e0 * 8); es.solar_gain = (double*)off(e->solar_gain, (size_t)e0 * 8);
    es.artificial_ratio = (const double*)off(e->artificial_ratio, (size_t)e0 * 8);
    es.max_power = (const double*)off(e->max_power, (size_t)e0 * 8); es.base_power = (double*)off(e->base_power, (size_t)e0 * 8);
    es.signal = (double*)off(e->signal, (size_t)e0 * 8); es.cluster_power = (double*)off(e->cluster_power, (size_t)e0 * 8);
    es.time_since_interp = (int32_t*)off(e->time_since_interp, (size_t)e0 * 4);
    es.perlin_seed = (const double*)off(e->perlin_seed, (size_t)e0 * 8);
    es.metrics = (double*)off(e->metrics, (size_t)e0 * MDR_N_METRICS * 8);
    // two slices are in flight at a time: each parity has its own due-tile queue in the workspace
    es.workspace = off(e->workspace, (size_t)(s & 1) * pipe_ws_bytes(cfg->n_envs));
    MdrStepInputs is = *in;
    is.actions = (const uint8_t*)off(in->actions, h0);
    is.od_noise = (const double*)off(in->od_noise, (size_t)e0 * 8);
    is.signal_noise = (const double*)off(in->signal_noise, (size_t)e0 * 8);
    is.interp_ids = (const int32_t*)off(in->interp_ids, (size_t)e0 * cfg->interp_nb_agents * 4);
    MdrOutputs os;
    os.obs = nullptr;
    os.reward = off(out->reward, h0 * rb);
    if (host_actions) {
      if (!in->actions) return MDR_ERR_NULL;
      CK(cudaMemcpyAsync(const_cast<uint8_t*>(is.actions), host_actions + h0, (size_t)ne * N, cudaMemcpyHostToDevice, st));
    }
    int status = run_steps_slice(&c, &hs, &es, &is, &os, e0, cfg->n_envs, st);
    if (status != MDR_OK) return status;
    void* dcomp = off(ctx->d_compact, h0 * 16 * rb);
    status = compact_slice(&c, &hs, &es, dcomp, st);
    if (status != MDR_OK) return status;
    CK(cudaMemcpyAsync(off(ctx->h_compact, h0 * 16 * rb), dcomp, (size_t)ne * N * 16 * rb, cudaMemcpyDeviceToHost, st));
    if (host_reward && out->reward)
      CK(cudaMemcpyAsync(off(host_reward, h0 * rb), os.reward, (size_t)ne * N * rb, cudaMemcpyDeviceToHost, st));
    if (host_power) CK(cudaMemcpyAsync(host_power + e0, es.cluster_power, (size_t)ne * 8, cudaMemcpyDeviceToHost, st));
    if (host_signal) CK(cudaMemcpyAsync(host_signal + e0, es.signal, (size_t)ne * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev_slice[s], st));
  }
  // the caller's stream continues after both internal streams
  for (int k = 0; k < 2; ++k) {
    CK(cudaEventRecord(ctx->ev_end[k], ctx->streams[k]));
    CK(cudaStreamWaitEvent(user_stream, ctx->ev_end[k], 0));
  }
  // expansion: every worker takes its share of every slice as soon as that slice has landed.  ONE thread (the caller's)
  // waits on the slice events and publishes the count of landed slices; the workers poll that counter.
  std::atomic<int> cuda_error{0};
  std::atomic<int> landed{0};
  const int T = ctx->n_threads;
  // chunks of a slice are claimed dynamically (a worker that lost its core for a moment does not hold up the slice)
  std::atomic<int> next_chunk[kMaxSlices];
  for (int s = 0; s < kMaxSlices; ++s) next_chunk[s].store(0, std::memory_order_relaxed);
  int chunk = (int)(262144 / ((size_t)N * (11 + 4 * C) * rb));  // ~256 KB of rows per claim
  if (chunk < 1) chunk = 1;
  auto expand_share = [&](int /*w*/, int s) {
    const int e0 = s * per, ne = (E - e0 < per ? E - e0 : per);
    thread_local std::vector<float> bf;
    thread_local std::vector<double> bd;
    for (;;) {
      const int a = next_chunk[s].fetch_add(chunk, std::memory_order_relaxed);
      if (a >= ne) break;
      const int b = a + chunk < ne ? a + chunk : ne;
      if (cfg->precision == MDR_F32) expand_envs<float>((const float*)ctx->h_compact, (float*)host_obs, e0 + a, e0 + b, N, C, bf);
      else expand_envs<double>((const double*)ctx->h_compact, (double*)host_obs, e0 + a, e0 + b, N, C, bd);
    }
  };
  auto wait_slices = [&]() {
    for (int s = 0; s < n_slices; ++s) {
      const cudaError_t err = cudaEventSynchronize(ctx->ev_slice[s]);
      if (err != cudaSuccess) { cuda_error.store((int)err); landed.store(n_slices + 1, std::memory_order_release); return; }
      landed.store(s + 1, std::memory_order_release);
    }
  };
  if (ctx->pool) {
    auto work = [&](int w) {
      for (int s = 0; s < n_slices; ++s) {
        while (landed.load(std::memory_order_acquire) <= s) _mm_pause();
        if (cuda_error.load() != 0) return;
        expand_share(w, s);
      }
    };
    ctx->pool->start(work);
    wait_slices();
    ctx->pool->finish();
  } else {
    for (int s = 0; s < n_slices; ++s) {
      const cudaError_t err = cudaEventSynchronize(ctx->ev_slice[s]);
      if (err != cudaSuccess) return fail(err);
      expand_share(0, s);
    }
  }
  if (cuda_error.load() != 0) return fail((cudaError_t)cuda_error.load());
  CK(cudaStreamSynchronize(user_stream));
#undef CK
  return MDR_OK;
}

}  // namespace mdr

extern "C" int mdr_host_ctx_create(const MdrConfig* cfg, int32_t n_threads, int32_t n_slices, MdrHostCtx** out_ctx) {
  if (!cfg || !out_ctx) return MDR_ERR_NULL;
  int st = mdr_validate(cfg);
  if (st != MDR_OK) return st;
  if (cudaSetDevice(cfg->device) != cudaSuccess) return MDR_ERR_CUDA;
  MdrHostCtx* c = new MdrHostCtx();
  c->device = cfg->device;
  if (n_threads <= 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    n_threads = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
    if (n_threads > 1) n_threads -= 1;  // the caller's thread issues the work and waits for the slices
    if (n_threads > 24) n_threads = 24;
  }
  if (n_threads < 1) n_threads = 1;
  c->n_threads = n_threads;
  c->n_slices = n_slices <= 0 ? 8 : (n_slices > kMaxSlices ? kMaxSlices : n_slices);
  c->compact_bytes = (size_t)cfg->n_envs * cfg->n_houses * 16 * (size_t)cfg->precision;
  bool ok = cudaStreamCreateWithFlags(&c->streams[0], cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->streams[1], cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_end[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_end[1], cudaEventDisableTiming) == cudaSuccess;
  // (blocking sync: the thread that waits for a slice sleeps instead of spinning on a core the expansion needs)
  for (int i = 0; ok && i < kMaxSlices; ++i)
    ok = cudaEventCreateWithFlags(&c->ev_slice[i], cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_compact, c->compact_bytes) == cudaSuccess;
  // pinned staging, first touched (and therefore placed) by this thread: pin the process to the GPU's NUMA node first
  ok = ok && cudaHostAlloc(&c->h_compact, c->compact_bytes, cudaHostAllocDefault) == cudaSuccess;
  if (!ok) {
    mdr_host_ctx_destroy(c);
    return MDR_ERR_CUDA;
  }
  memset(c->h_compact, 0, c->compact_bytes);
  if (n_threads > 1) c->pool = new Pool(n_threads);
  *out_ctx = c;
  return MDR_OK;
}

extern "C" int mdr_host_ctx_destroy(MdrHostCtx* c) {
  if (!c) return MDR_OK;
  cudaSetDevice(c->device);
  delete c->pool;
  for (int k = 0; k < 2; ++k) {
    if (c->streams[k]) cudaStreamDestroy(c->streams[k]);
    if (c->ev_end[k]) cudaEventDestroy(c->ev_end[k]);
  }
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  for (int i = 0; i < kMaxSlices; ++i)
    if (c->ev_slice[i]) cudaEventDestroy(c->ev_slice[i]);
  if (c->d_compact) cudaFree(c->d_compact);
  if (c->h_compact) cudaFreeHost(c->h_compact);
  delete c;
  return MDR_OK;
}

extern "C" int mdr_host_ctx_info(const MdrHostCtx* c, int32_t* n_threads, int32_t* n_slices, size_t* compact_bytes) {
  if (!c) return MDR_ERR_NULL;
  if (n_threads) *n_threads = c->n_threads;
  if (n_slices) *n_slices = c->n_slices;
  if (compact_bytes) *compact_bytes = c->compact_bytes;
  return MDR_OK;
}

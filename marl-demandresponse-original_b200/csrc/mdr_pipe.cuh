// Persistent software-pipelined step kernel (fp32 fast path), its prologue warp and deferred interpolation refresh.
// Included by mdr_kernels.cu inside namespace mdr (after the shared device helpers); not a standalone translation unit.
#pragma once

// ----------------------------------------------------------------------------------------
// Persistent, software-pipelined variant of the fast path (fp32, default observation layout; solar gain,
// message drops and the metric accumulators are variants of it).  The CTAs stay resident (SMs x CTAs/SM
// of them) and loop over tiles of G envs.  Two schedules (template parameter kDyn): a fixed, strided list
// of tiles per CTA, described first, and tiles claimed in address order (large problems; see "In-order
// tile claiming" below).  With the strided lists:
//   * the house threads' inputs of tile i+1 are fetched with cp.async into a second shared-memory
//     stage while tile i is computed (each thread copies and later reads only its own record, so
//     cp.async.wait_group is the only synchronisation the inputs need);
//   * the dedicated prologue warp runs `pro_batch` tiles per pass and up to 2*pro_batch tiles ahead,
//     handing a 64-byte PipeEnv record per env over through an mbarrier ring, and writes the
//     per-env outputs (clock, outdoor temperature, signal) itself;
//   * the bulk (TMA) observation store of tile i drains while tile i+1 is loaded and updated.
// Barrier 1 = house warps only (message window + power partial sums); everything that is constant
// over the tile loop (shared-memory addresses, neighbour window, partial-sum slots) is computed
// once per thread before the loop.
// ----------------------------------------------------------------------------------------
// Trace build (-DMDR_TRACE): lane 0 of every warp of CTA MDR_TRACE_CTA stamps globaltimer at fixed
// points of its first MDR_TRACE_TILES tiles; tools/trace_tile.py reads the buffer back.
#ifdef MDR_TRACE
#ifndef MDR_TRACE_CTA
#define MDR_TRACE_CTA 200
#endif
#define MDR_TRACE_TILES 24
#define MDR_TRACE_POINTS 10
__device__ unsigned long long g_trace[8 * MDR_TRACE_TILES * MDR_TRACE_POINTS];
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define MDR_STAMP_AT(w, t, k)                                                                        \
  do {                                                                                                \
    if (blockIdx.x == MDR_TRACE_CTA && (threadIdx.x & 31) == 0 && (t) < MDR_TRACE_TILES)              \
      g_trace[((w) * MDR_TRACE_TILES + (t)) * MDR_TRACE_POINTS + (k)] = global_ns();                  \
  } while (0)
#define MDR_STAMP(k) MDR_STAMP_AT(warp, it, k)
__device__ void trace_stamp_fn(int w, int t, int k) { MDR_STAMP_AT(w, t, k); }
// per-CTA span of the launch: [0] kernel entry, [1] past griddepcontrol.wait, [2] tile loop done, [3] kernel exit
__device__ unsigned long long g_cta_span[2048 * 8];
#define MDR_CTA_STAMP(k)                                                                              \
  do {                                                                                                \
    if (threadIdx.x == 0 && blockIdx.x < 2048) g_cta_span[blockIdx.x * 8 + (k)] = global_ns();        \
  } while (0)
}  // namespace mdr
extern "C" int mdr_debug_cta_span(unsigned long long* host, size_t n) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -5;
  if (cudaMemcpyFromSymbol(host, mdr::g_cta_span, n * sizeof(unsigned long long)) != cudaSuccess) return -5;
  return 0;
}
// trace builds only (tools/trace_tile.py): copies the globaltimer stamps of the traced CTA to the host
extern "C" int mdr_debug_trace(unsigned long long* host, size_t n, int clear) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -5;
  if (cudaMemcpyFromSymbol(host, mdr::g_trace, n * sizeof(unsigned long long)) != cudaSuccess) return -5;
  if (clear) {
    void* ptr = nullptr;
    cudaGetSymbolAddress(&ptr, mdr::g_trace);
    cudaMemset(ptr, 0, n * sizeof(unsigned long long));
  }
  return 0;
}
namespace mdr {
#else
#define MDR_STAMP_AT(w, t, k) do { } while (0)
#define MDR_STAMP(k) do { } while (0)
#define MDR_CTA_STAMP(k) do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_16(void* s, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* s, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* s, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// mbarrier hand-over between the prologue warp (producer of PipeEnv records) and the house warps
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// shared-memory control block of the pipelined kernel (at off_ctl).  The PipeEnv ring has
// ring = 2 * pro_batch <= kMaxRing slots; slot = it % ring for the it-th tile of this CTA.
constexpr int kMaxRing = 16;
constexpr int kMaxDue = 60;
struct PipeCtl {
  uint64_t full[kMaxRing];   // prologue -> house warps: slot is ready            (count 1)
  uint64_t empty[kMaxRing];  // house warps -> prologue: slot may be overwritten   (count house_warps)
  int tile_due[kMaxRing];    // any env of the tile has an interpolation refresh due
  // tiles of this CTA with a refresh due in this launch: with staggered refresh clocks (rollouts restart clusters at
  // different times) a few percent of the tiles are due at EVERY step, and the deferred pass must only visit those
  int due_n;
  int due_list[kMaxDue];
  int pub_count[2];  // warps that have fenced a due tile of even / odd position (shared due queue: early publication)
  int ring_tile[4];  // in-order tile claiming: tile of position it in slot it & 3 (kDuePhase flag), written by the claim warp
};

// The prologue warp produces `pro_batch` tiles per pass: the 32 lanes are split into pro_batch
// groups (one per tile), each group into lane sets of `pro_lanes` lanes per env.  The deeper the
// batch, the further the prologue's dependent fp64 / Philox / global-load chains are from the
// house warps' critical path.
// One pass: tiles it0 .. it0+B-1 of this CTA (B a power of two <= pro_batch).
__device__ __noinline__ void prologue_pass(const KernelParams& p, int it0, int B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PipeEnv* s_env = reinterpret_cast<PipeEnv*>(smem_raw + p.off_env);
  PipeCtl& ctl = *reinterpret_cast<PipeCtl*>(smem_raw + p.off_ctl);
  const int ring = 2 * p.pro_batch;
  const int ring_shift = 31 - __clz(ring);
  const int lane = threadIdx.x & 31;
  const int lanes_per_tile = 32 / B;
  int L = 16;  // lanes cooperating on one env (nb_octaves perlin octaves + 1 Philox normal per env)
  while (L > 1 && L * p.G > lanes_per_tile) L >>= 1;
  const int groups = lanes_per_tile / L;  // envs of a tile processed at once (>= 1: G * pro_batch <= 32)
  const int tlane = lane & (lanes_per_tile - 1);
  const int sub = tlane & (L - 1), grp = tlane / L;
  const int my_k = lane / lanes_per_tile;  // which tile of the pass this lane works for
  const unsigned group_mask = (lanes_per_tile == 32 ? 0xffffffffu : ((1u << lanes_per_tile) - 1u)) << (my_k * lanes_per_tile);
  const int it = it0 + my_k;
  const int tile = blockIdx.x + it * gridDim.x;
  const bool tile_valid = tile < p.n_tiles;
  const int slot = it & (ring - 1);
  MDR_STAMP_AT(7, it0, 8);
  // the house warps must have released the pass's ring slots.  Warp-uniform loop over the B barriers:
  // per-lane-group waits on different mbarriers (a divergent try_wait spin) were measured to return
  // up to 8 us late (tools/trace_tile.py)
  for (int k = 0; k < B; ++k) {
    const int itk = it0 + k;
    if (blockIdx.x + itk * gridDim.x < p.n_tiles && (itk >> ring_shift) >= 1)
      mbar_wait(&ctl.empty[itk & (ring - 1)], (((itk >> ring_shift) & 1) ^ 1));
  }
  MDR_STAMP_AT(7, it0, 0);
  const int tile_c = tile_valid ? tile : blockIdx.x;  // lanes of an absent tile compute but never write
  // (split env, mdr_pipe_split.cuh: the tile is the slice of rank tile % cl of env tile / cl)
  const int env0 = p.cl > 1 ? tile_c / p.cl : tile_c * p.G;
  const int genvs = p.cl > 1 ? 1 : min(p.G, p.E - env0);
  PipeEnv* buf = s_env + slot * p.G;
  EnvScratch unused;
  int my_due = 0;
  for (int first = 0; first < p.G; first += groups) {  // warp-uniform trip count (G, not genvs)
    const int le2 = first + grp;
    const bool valid = tile_valid && le2 < genvs;
    const int lec = le2 < genvs ? le2 : genvs - 1;
    my_due |= env_prologue<true>(p, unused, buf[lec], env0 + lec, sub, L, valid, false, false);
  }
  const unsigned due_ballot = __ballot_sync(0xffffffffu, my_due != 0);
  if (tile_valid && tlane == 0) ctl.tile_due[slot] = (due_ballot & group_mask) != 0;
  __syncwarp();
  if (tile_valid && tlane == 0) mbar_arrive(&ctl.full[slot]);
  MDR_STAMP_AT(7, it0, 7);
}

// Queue of tiles with an interpolation refresh due, shared by ALL CTAs of a launch (in MdrEnvs.workspace, zeroed by its
// owner once; the kernel leaves it zeroed).  With staggered refresh clocks a few percent of the tiles are due at every
// step: a CTA that refreshed its own due tiles after its tile loop would hold the whole launch back by 10+ us per
// tile (measured: 30-40 us of tail on 16 384 x 100), so every CTA publishes its due tiles here and then takes tiles
// from the queue until it is empty -- the refresh work of a step is spread over the whole grid.
struct DueQueue {
  unsigned reserved;   // slots handed out to publishers
  unsigned taken;      // slots claimed by consumers
  unsigned ctas_done;  // CTAs that have published everything they have
  unsigned exited;     // CTAs that are done with the queue (the last one re-zeroes the header)
  unsigned pad[12];
  unsigned tiles[1];   // [n_tiles] tile index + 1, 0 = empty
};

// ----------------------------------------------------------------------------------------
// In-order tile claiming (KernelParams.dyn_off != 0).  With a fixed, strided tile list per CTA the CTAs of a launch
// drift apart (per-CTA trace of 16 384 x 100: tile loops end between 67 and 84 us) -- the launch waits for the slowest
// one, and the drifting write front costs DRAM locality (tools/microbench/write_patterns_bw.cu: 5.85 TB/s against
// 6.85 TB/s for tiles claimed in address order).  Claiming needs the per-env record of a tile that nobody planned to
// process in this CTA, so the hand-over leaves shared memory:
//   * at the start of the launch ALL warps of the grid (house warps and the prologue warp alike) produce the 64-byte
//     records of all envs into a global (L2-resident) array -- pro_batch tiles per warp, the same env_prologue code --
//     plus a `due` word per tile and the launch's list of due tiles, and meet at a grid barrier (one arrival per CTA);
//   * the eighth warp of the CTA then is its CLAIM WARP: it maps positions of the launch's work list to tiles (the
//     first three of a CTA are fixed, the others come from an atomic counter, claimed three tiles ahead), and for every
//     position puts the tile id, the tile's `due` word and its records into a 4-slot ring (cp.async + mbarrier
//     complete); the house warps wait for a slot one tile before they need it -- no global round trip on their path;
//   * due tiles come first in the work list and are refreshed by the CTA that processed them, between two tiles
//     (pipe_refresh_own); a due tile met again in address order is skipped.
// Measured alternatives, not kept (16 384 x 100, in-phase refresh clocks, static lists 90.0 us per step; final: 86.4):
//   * records from the CTA's own prologue warp (strided slice, no back-pressure) with a ready flag per tile: every
//     global round trip the house warps depend on (flag, then record) needs a whole tile of slack under the observation
//     write stream (loaded L2 latency 1-2 us; acquire loads also invalidate L1): 95-116 us;
//   * a small kernel in front of the step kernel producing all records (512 CTAs, 6.5 us) -- step kernel 73 us, but
//     13 us from one step kernel to the next: 86.1 us;
//   * thread 0 of the house warps claiming and a few loader threads copying the records: the claim logic and the
//     refresh call inside the tile loop cost the loop registers (see pipe_refresh_own).
// ----------------------------------------------------------------------------------------
struct DynHdr {
  unsigned next_tile;  // positions handed out beyond the first three of every CTA
  unsigned exited;     // CTAs that are done claiming (the last one out re-zeroes the header)
  unsigned arrived;    // CTAs that have written their share of the records (grid barrier in front of the tile loop)
  unsigned n_due;      // tiles with an interpolation refresh due in this launch = entries of the due list
  unsigned pad[12];    // (arrived | n_due are read with one 8-byte load: n_due is final once every CTA has arrived)
};
constexpr int kDuePhase = 1 << 30;  // flag on a claimed tile index: taken from the due list (processed AND refreshed by this CTA)
static_assert(sizeof(DynHdr) == 64, "DynHdr must stay 64 bytes");

__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ DynHdr* dyn_hdr(const KernelParams& p) {
  return reinterpret_cast<DynHdr*>(reinterpret_cast<unsigned char*>(p.workspace) + p.dyn_off);
}
__device__ __forceinline__ unsigned* dyn_due(const KernelParams& p) { return reinterpret_cast<unsigned*>(dyn_hdr(p) + 1); }
__device__ __forceinline__ unsigned* dyn_list(const KernelParams& p) {
  return reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(p.workspace) + p.dyn_list_off);
}
__device__ __forceinline__ PipeEnv* dyn_recs(const KernelParams& p) {
  return reinterpret_cast<PipeEnv*>(reinterpret_cast<unsigned char*>(p.workspace) + p.dyn_rec_off);
}

// Per-env prologue of a whole step by ALL warps of the step kernel (pro_batch tiles per warp and pass): records into
// the global array, per-env outputs in place (as the strided lists' prologue warp does), a `due` word per tile and the
// launch's due-tile count.  `warp_g` / `n_warps`: this warp's index among / the number of warps of the grid.
// (Out of line.  Its timeline in a trace build, row 22: 1.2 us until the per-env loads are back, 1.7 us of draws --
//  including a second round trip for the perlin seed; loading it with the others just moved the time -- 0.6 us sinpi,
//  0.3 us signal, 1.7 us until the record / in-place stores and the due word are out.  Inlining it into the kernel
//  changed nothing of that and spilled 150 bytes in the kernel.)
__device__ __noinline__ void pro_all_tiles(const KernelParams& p, int warp_g, int n_warps) {
  const int lane = threadIdx.x & 31;
  // B tiles per warp and pass (pro_batch, a power of two with B * G <= 32), L lanes per env
  const int B = p.pro_batch;
  const int lanes_per_tile = 32 / B;
  int L = 16;
  while (L > 1 && L * p.G > lanes_per_tile) L >>= 1;
  const int groups = lanes_per_tile / L;  // envs of a tile processed at once
  const int tlane = lane & (lanes_per_tile - 1);
  const int sub = tlane & (L - 1), grp = tlane / L;
  const int my_k = lane / lanes_per_tile;
  const unsigned group_mask = (lanes_per_tile == 32 ? 0xffffffffu : ((1u << lanes_per_tile) - 1u)) << (my_k * lanes_per_tile);
  PipeEnv* const recs = dyn_recs(p);
  MDR_STAMP_AT(threadIdx.x >> 5, 22, 0);
  for (int tile0 = warp_g * B; tile0 < p.n_tiles; tile0 += n_warps * B) {
  const int tile = tile0 + my_k;
  const bool tile_valid = tile < p.n_tiles;
  const int tile_c = tile_valid ? tile : 0;  // lanes of an absent tile compute but never write
  const int env0 = tile_c * p.G;
  const int genvs = min(p.G, p.E - env0);
  EnvScratch unused;
  int my_due = 0;
  for (int first = 0; first < p.G; first += groups) {  // warp-uniform trip count
    const int le2 = first + grp;
    const bool valid = tile_valid && le2 < genvs;
    const int lec = le2 < genvs ? le2 : genvs - 1;
    // (fields stored one by one: a local record costs this function sixteen registers it does not have -- it is compiled
    //  under the step kernel's 80-register cap, and its spills are L2 round trips here)
    my_due |= env_prologue<true>(p, unused, recs[env0 + lec], env0 + lec, sub, L, valid, false, false);
  }
  const unsigned due_ballot = __ballot_sync(0xffffffffu, my_due != 0);
  if (tile_valid && tlane == 0) {
    const bool due = (due_ballot & group_mask) != 0;
    __stcg(dyn_due(p) + tile, due ? 1u : 0u);
    if (due) __stcg(dyn_list(p) + atomicAdd(&dyn_hdr(p)->n_due, 1u), (unsigned)tile);  // the launch's due list
  }
  }
  MDR_STAMP_AT(threadIdx.x >> 5, 22, 6);
}

// Deferred interpolation refresh (every interp_update_period seconds; PowerGrid.step :1250-1255,
// interpolatePower :1195-1234).  It runs on 1 step in 75, needs fp64 and a 32-corner table walk per
// house, and would cost the tile loop registers if it sat inside it.  So the tile loop treats a due
// env like any other (the prologue parks its perlin value and marks it), and this pass -- after the
// loop, same launch, same CTA, same tile order -- evaluates the table on the houses' NEW state,
// re-evaluates the signal and patches observation feature 9 (the only output that depends on it).
// signal-dependent metric accumulators of one env and step (MDR_M_SUM_SIGNAL ..., main-deploy.py:140-149)
// (Accumulators are advanced with reductions at L2 -- RED.ADD.F64, no value comes back: a load-add-store by the env's
//  first thread put a 1-2 us global round trip per tile on that warp's critical path, 126 instead of 92 us per c4 step.
//  One contributor per env and launch, launches in stream order: the sums are the same bits as a read-modify-write.)
__device__ __forceinline__ void metrics_signal_terms(double* m, double sig, double P) {
  const double d = sig - P;
  atomicAdd(m + MDR_M_SUM_SIGNAL, sig);
  atomicAdd(m + MDR_M_SUM_SIGNAL_OFFSET, d);
  atomicAdd(m + MDR_M_SUM_SIGNAL_ERROR, fabs(d));
  atomicAdd(m + MDR_M_SUM_SQ_SIGNAL_ERROR, d * d);
}


// refresh of ONE due tile (all house threads of the CTA): table walk on the houses' NEW state, base power, signal,
// observation feature 9 (PowerGrid.step :1250-1255, interpolatePower :1195-1234)
template <bool kOwn = false>  // kOwn: the CTA's own tile, its bulk row stores may still be in flight (waited for in front of the patch)
__device__ __forceinline__ void pipe_refresh_tile(const KernelParams& p, int tile, int le, int li) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_val = reinterpret_cast<double*>(smem_raw + p.off_val);
  float* s_fsig = reinterpret_cast<float*>(smem_raw + p.off_pw);  // [G], the power partials are dead by now
  const int tid = threadIdx.x;
  const int N = p.N, G = p.G, GN = G * N;
  const int nb = p.interp_nb_agents;
  const int nsamp = N <= nb ? N : nb;
  const int T = p.hmax;
  const int H = min(GN, (p.E - tile * G) * N);
  const bool active = tid < H;
  const int e = tile * G + le;
  const unsigned h = (unsigned)tile * (unsigned)GN + (unsigned)tid;
  MDR_STAMP_AT(tid >> 5, MDR_TRACE_TILES - 1, 0);  // (trace builds: the refresh's own timeline in the last row)
  // Everything that depends on nothing else is loaded up front, in flight together: under the observation write stream
  // a dependent global round trip costs 2-3 us, and the refresh is nothing but a chain of them (it took 18 us per tile
  // with the loads issued where their values were needed).
  // (another SM may have written this tile's state in this launch: L1 is not coherent across SMs, read through L2)
  const int tsi = active ? __ldcg(p.time_since_interp + e) : 0;
  double od_new = active ? __ldcg(p.od_temp + e) : 0.0;
  const bool head = active && li == 0;
  const bool need_clock = head || (active && p.solar);
  const uint32_t t_ep = need_clock ? (uint32_t)__ldcg(p.t_epoch + e) : 0u;
  // (the env's first thread parks what it needs after the table walk in shared memory: registers are what this function
  //  lacks -- it runs under the step kernel's 80-register cap and a spill is an L2 round trip here)
  double* const s_head = s_val + T + G * 32 + le * 8;
  if (head) {
    s_head[0] = __ldcg(p.base_power + e);  // perlin value parked by the prologue
    s_head[1] = p.artificial_ratio[e];
    s_head[2] = p.max_power[e];
    s_head[3] = p.metrics != nullptr ? __ldcg(p.cluster_power + e) : 0.0;
    s_head[4] = (double)t_ep;
  }
  const bool direct = N <= nb;  // every house is sampled: thread li evaluates house li
  float2 t2 = make_float2(0.f, 0.f);
  double tg = 0.0;
  int key = 0;
  if (active && direct) {
    const size_t hs = (size_t)e * N + li;
    t2 = __ldcg(reinterpret_cast<const float2*>(p.temps) + hs);
    tg = (double)reinterpret_cast<const float4*>(p.coef_b)[hs].w;
    key = p.interp_key[hs];
  }
  const bool due = active && tsi < 0;
  double hour_s = 0.0, date = 0.0;
  if (due && p.solar) {  // interpolatePower point :1198-1207: seconds since midnight and tm_yday, 0 with solar gain off
    Calendar cal = calendar_time(t_ep);
    calendar_date(cal);
    hour_s = (double)cal.sod;
    date = (double)cal.yday;
  }
  double val = 0.0;
  if (due && li < nsamp) {
    if (!direct) {
      int src;
      if (p.interp_ids) src = p.interp_ids[(size_t)e * nb + li];
      else {
        const uint4 r = philox4x32((uint32_t)(e + p.env_base), (uint32_t)step_now(p), (uint32_t)(step_now(p) >> 32),
                                   STREAM_IDS + 16 * (uint32_t)li, p.seed);
        src = (int)(((uint64_t)r.x * (uint64_t)N) >> 32);
      }
      const size_t hs = (size_t)e * N + src;
      t2 = __ldcg(reinterpret_cast<const float2*>(p.temps) + hs);
      tg = (double)reinterpret_cast<const float4*>(p.coef_b)[hs].w;
      key = p.interp_key[hs];
    }
    val = interp_eval_grid<float, InterpGrid>(*reinterpret_cast<const InterpGrid*>(smem_raw + p.off_grid), p.interp_table,
                                              key, (double)t2.x - tg, (double)t2.y - tg, od_new - tg, hour_s, date);
  }
  MDR_STAMP_AT(tid >> 5, MDR_TRACE_TILES - 1, 1);
  s_val[tid] = val;  // 0 for houses that are not sampled
  house_sync(T);
  MDR_STAMP_AT(tid >> 5, MDR_TRACE_TILES - 1, 2);
  // per-env sum by the env's first 32 threads (the fp32 mode's tolerance does not need the reference's id order here --
  // the fp64 kernels keep it)
  if (due && li < 32) {
    double part = 0.0;
    for (int i = li; i < nsamp; i += 32) part += s_val[le * N + i];
    s_val[T + le * 32 + li] = part;
  }
  house_sync(T);
  if (due && li == 0) {
    double base = 0.0;
    const int np = nsamp < 32 ? nsamp : 32;
    for (int i = 0; i < np; ++i) base += s_val[T + le * 32 + i];
    if (N > nb) base = mul_rn(base, (double)N / (double)nb);
    const Calendar cal = calendar_time((uint32_t)s_head[4]);
    const int time_sec = cal.hour * 3600 + cal.minute * 60 + cal.second;
    const double sig = grid_signal(p, base, time_sec, s_head[0], s_head[1], s_head[2]);
    p.base_power[e] = base;
    p.time_since_interp[e] = 0;
    p.signal[e] = sig;
    s_fsig[le] = (float)(sig * p.inv_norm_sig_agents);
    // the tile loop left the signal-dependent accumulators of a due env to this pass (its signal was not final)
    if (p.metrics != nullptr) metrics_signal_terms(p.metrics + (size_t)e * MDR_N_METRICS, sig, s_head[3]);
  }
  MDR_STAMP_AT(tid >> 5, MDR_TRACE_TILES - 1, 3);
  if (kOwn && (tid & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every warp's rows have landed ...
  MDR_STAMP_AT(tid >> 5, MDR_TRACE_TILES - 1, 4);
  house_sync(T);                                                                           // ... before anybody patches one
  MDR_STAMP_AT(tid >> 5, MDR_TRACE_TILES - 1, 5);
  if (due && p.obs != nullptr) reinterpret_cast<float*>(p.obs)[(size_t)h * p.F + 9] = s_fsig[le];
  // (s_val / s_fsig are reused by the next tile: its first writes come after this tile's last reads of the same
  //  thread's entries, and its first barrier orders the rest)
}

// Publication of a due tile to the shared queue, one tile after it was processed (called by every house thread): once
// every warp's bulk store of that tile has completed and its state stores are fenced, the last warp hands the tile to
// whichever CTA is idle first -- the refresh of a due tile overlaps the rest of its owner's tile loop.
__device__ __noinline__ void publish_due_tile(DueQueue* q, int tile, int* counter, int house_warps) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __threadfence();
  __syncwarp();
  if (lane == 0) {
    if (atomicAdd(counter, 1) == house_warps - 1) {
      *counter = 0;
      __threadfence();
      const unsigned idx = atomicAdd(&q->reserved, 1u);
      atomicExch(&q->tiles[idx], (unsigned)tile + 1u);
    }
  }
}

// Deferred interpolation refresh (every interp_update_period seconds).  It runs on 1 step in 75 per env, needs fp64
// and a 32-corner table walk per house, and would cost the tile loop registers if it sat inside it.  So the tile loop
// treats a due env like any other (the prologue parks its perlin value and marks it), and this pass -- after the loop,
// same launch -- evaluates the table on the houses' NEW state, re-evaluates the signal and patches observation
// feature 9 (the only output that depends on it).
__device__ __noinline__ void pipe_refresh_pass(const KernelParams& p, int le, int li, int pend_tile) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int T = p.hmax;
  PipeCtl& ctl = *reinterpret_cast<PipeCtl*>(smem_raw + p.off_ctl);
  // this CTA's state stores and the bulk stores of its rows must have landed (and be visible to the whole GPU: another
  // CTA may refresh these tiles) before anything is published or patched
  if ((tid & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __threadfence();
  house_sync(T);  // ... and the due-tile list written by thread 0 during the tile loop is complete
  DueQueue* q = reinterpret_cast<DueQueue*>(p.workspace);
  if (q == nullptr) {
    // no shared queue (no workspace): this CTA refreshes its own due tiles (listed by thread 0 during the loop; when
    // the list overflowed, every tile of the CTA is visited)
    const int n_due = ctl.due_n;
    const bool listed = n_due <= kMaxDue;
    const int n_visit = listed ? n_due : (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    for (int v = 0; v < n_visit; ++v)
      pipe_refresh_tile(p, listed ? ctl.due_list[v] : (int)blockIdx.x + v * (int)gridDim.x, le, li);
    return;
  }
  int* s_take = &ctl.tile_due[0];  // (the ring's due flags are dead after the tile loop)
  if (tid == 0) {
    if (pend_tile >= 0) {  // the CTA's last tile was due (the earlier ones were published during the loop)
      const unsigned idx = atomicAdd(&q->reserved, 1u);
      atomicExch(&q->tiles[idx], (unsigned)pend_tile + 1u);
    }
    __threadfence();
    atomicAdd(&q->ctas_done, 1u);
  }
  for (;;) {
    if (tid == 0) {
      const unsigned idx = atomicAdd(&q->taken, 1u);
      int got = -1;
      unsigned long long t0 = 0;
      for (unsigned spin = 0;; ++spin) {
        const unsigned v = ld_volatile_u32(&q->tiles[idx]);
        if (v != 0) {
          q->tiles[idx] = 0;  // leave the queue zeroed for the next launch
          got = (int)v - 1;
          break;
        }
        if (ld_volatile_u32(&q->ctas_done) == gridDim.x) {  // every publisher is done: reserved is final
          if (idx >= ld_volatile_u32(&q->reserved)) break;
          continue;  // the slot was written before its publisher counted itself done: re-read it
        }
        if ((spin & 1023) == 1023) {  // a stuck peer must not hang the GPU: give up after ~2 s
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          else if (now - t0 > 2000000000ull) break;
        }
      }
      __threadfence();  // (acquire side of the publisher's fence)
      *s_take = got;
    }
    house_sync(T);
    const int tile = *s_take;
    if (tile < 0) break;
    pipe_refresh_tile(p, tile, le, li);
    house_sync(T);  // s_take is rewritten by thread 0
  }
  if (tid == 0) {
    if (atomicAdd(&q->exited, 1u) == gridDim.x - 1) {  // last one out re-zeroes the header
      q->reserved = 0; q->taken = 0; q->ctas_done = 0; q->exited = 0;
      __threadfence();
    }
  }
}

// In-order claiming: a tile taken from the launch's due list is refreshed by the CTA that processed it, one tile later
// (inside the tile loop: the other CTAs simply claim more tiles meanwhile, so the refresh work of a step costs the launch
// its share of the grid's time instead of a tail).  Every warp's bulk stores of the tile's rows must have landed before
// feature 9 is patched; the tile's state was stored before the previous tile barrier of this CTA.
// The tile loop LEAVES for this call and is re-entered afterwards (its thirty-odd loop-invariant registers are recomputed):
// a call inside the loop made the compiler spill some of them for good, and with 3 x 76 KB of shared memory per SM
// there is no L1 left for local memory -- every spill reload was an L2 round trip (16 384 x 100: 127 instead of 88 us).
__device__ __noinline__ void pipe_refresh_own(const KernelParams& p, int tile) {
  const int tid = threadIdx.x;
  const int le = tid < p.G * p.N ? (p.N == 1 ? tid : (int)__umulhi((unsigned)tid, p.div_magic)) : 0;  // tid / N
  const int li = tid - le * p.N;
  // (the tile's state was stored by this CTA's threads in front of at least one house-warp barrier; the bulk stores of
  //  its rows are waited for inside, in front of the patch, so that they drain under the table walk)
  pipe_refresh_tile<true>(p, tile, le, li);
}

// kVar: bit 0 = metric accumulators (MDR_M_*) as an epilogue of the tile; bit 1 = message drops
// (replayed msg_keep, or Philox against comm_defect_prob) in the observation rows.
template <int kC, int kAct, bool kObs, int kVar, bool kDyn>
__global__ void __launch_bounds__(256, 3) step_pipe_kernel(const __grid_constant__ KernelParams p) {
  constexpr bool kMetrics = (kVar & 1) != 0;
  constexpr bool kDrops = kObs && (kVar & 2) != 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  PipeCtl& ctl = *reinterpret_cast<PipeCtl*>(smem_raw + p.off_ctl);
  constexpr bool dyn = kDyn;  // in-order tile claiming (see DynHdr): 4-slot ring of tile ids + records filled by the claim warp
  const int ring_mask = dyn ? 3 : 2 * p.pro_batch - 1;
  const int ring_shift = 31 - __clz(ring_mask + 1);
  MDR_CTA_STAMP(0);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the next step may start occupying freed SMs
  if (p.base_power_mode == MDR_BASE_INTERPOLATION) {  // shared-memory copy of the interpolation grid (see InterpGrid)
    InterpGrid* g = reinterpret_cast<InterpGrid*>(smem_raw + p.off_grid);
    if (tid < MDR_INTERP_DIMS) g->interp_dims[tid] = p.interp_dims[tid];
    for (int i = tid; i < MDR_INTERP_DIMS * MDR_INTERP_MAX_AXIS; i += blockDim.x)  // (a CTA may have fewer than 120 threads)
      (&g->interp_axes[0][0])[i] = (&p.interp_axes[0][0])[i];
  }
  if (tid == 0) {
    ctl.due_n = 0;
    ctl.pub_count[0] = ctl.pub_count[1] = 0;
    for (int i = 0; i <= ring_mask; ++i) {
      mbar_init(&ctl.full[i], dyn ? 33 : 1);  // (claim warp: 32 copy completions + lane 0's release of the tile id)
      mbar_init(&ctl.empty[i], p.house_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (dyn) {
    // Every warp of the grid (house warps and the prologue warp alike) produces its share of the step's per-env
    // records, then the CTA arrives at the grid barrier.  The CTAs of a launch are co-resident (grid <= SMs x resident
    // CTAs; a dependent launch only takes the slots this one frees).
    asm volatile("griddepcontrol.wait;" ::: "memory");  // everything the previous launch wrote is visible from here on
    pro_all_tiles(p, warp * gridDim.x + blockIdx.x, (blockDim.x >> 5) * gridDim.x);
    cta_sync();
    if (warp >= p.house_warps) {
      // ---- the claim warp: grid barrier, then the ring of (tile id, due word, records) the house warps consume ----
      // Position q of the launch's work list is tile q for q < grid, entry q - grid of the due list (kDuePhase) for the
      // next n_due positions, tile q - n_due after that.  The first three positions of a CTA are fixed (blockIdx + 0, 1,
      // 2 x grid), the others come from the counter -- position it + 3 is claimed when every house warp has started
      // tile it.  A due tile met in address order is skipped by the house warps.
      DynHdr* const dh = dyn_hdr(p);
      unsigned D = 0;
      if (lane == 0) {
        __threadfence();
        atomicAdd(&dh->arrived, 1u);
        unsigned long long t0 = 0;
        unsigned long long both;  // arrived | n_due << 32, one 8-byte load: n_due is final once every CTA has arrived
        for (unsigned spin = 1;; ++spin) {
          asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(both) : "l"(&dh->arrived) : "memory");
          if ((unsigned)both >= gridDim.x) break;
          __nanosleep(32);
          if ((spin & 1023u) == 0) {  // a CTA that never becomes resident must not hang the GPU: abort loudly after ~2 s
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
          }
        }
        __threadfence();
        D = (unsigned)(both >> 32);
      }
      D = __shfl_sync(0xffffffffu, D, 0);
      const int n_tiles = p.n_tiles;
      auto map_position = [&](unsigned q) -> int {  // (lane 0)
        if (q < gridDim.x) return (int)q;
        if (q < gridDim.x + D) return (int)ld_volatile_u32(dyn_list(p) + (q - gridDim.x)) | kDuePhase;
        const unsigned t = q - D;
        return t < (unsigned)n_tiles ? (int)t : n_tiles;
      };
      auto fill = [&](int slot, int enc) {  // tile id, due word and records of one position into ring slot `slot`
        const int t = enc & (kDuePhase - 1);
        if (t < n_tiles) {
          if (lane == 0) cp_async_4(&ctl.tile_due[slot], dyn_due(p) + t);
          const int nchunks = 4 * min(p.G, p.E - t * p.G);
          unsigned char* dst = smem_raw + p.off_env + slot * p.G * (int)sizeof(PipeEnv);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(dyn_recs(p) + (size_t)t * p.G);
          for (int c = lane; c < nchunks; c += 32) cp_async_16(dst + c * 16, src + c * 16);
        }
        cp_async_mbar_arrive_noinc(&ctl.full[slot]);
        if (lane == 0) {
          ctl.ring_tile[slot] = enc;
          mbar_arrive(&ctl.full[slot]);
        }
      };
      int enc = 0;
      fill(0, (int)blockIdx.x);
      for (int j = 1; j < 3; ++j) {
        if (lane == 0) enc = map_position(j * gridDim.x + blockIdx.x);
        enc = __shfl_sync(0xffffffffu, enc, 0);
        fill(j, enc);
      }
      for (int j = 3; (enc & (kDuePhase - 1)) < n_tiles; ++j) {
        const int slot = j & 3;
        // permission: every house warp has started tile j - 3 (and is done with the slot's previous tenant, tile j - 4).
        // The claim (an atomic round trip) and the record copies (another one) then have two tiles to land: under the
        // observation write stream each takes 1-2 us, and with one tile of slack the house warps waited for them.
        mbar_wait(&ctl.empty[slot], ((j >> 2) - (slot < 3 ? 1 : 0)) & 1);
        if (lane == 0) enc = map_position(3u * gridDim.x + atomicAdd(&dh->next_tile, 1u));
        enc = __shfl_sync(0xffffffffu, enc, 0);
        fill(slot, enc);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      return;
    }
  }
  if (warp >= p.house_warps) {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // everything the previous launch wrote is visible from here on
    // tiles per pass: pro_batch, but not more than this launch gives a CTA (a small problem should spend its lanes
    // on the envs of tiles that exist, not on absent ones)
    int B = p.pro_batch;
    while (B > 1 && (B >> 1) * (int)gridDim.x >= p.n_tiles) B >>= 1;
    for (int it0 = 0; blockIdx.x + it0 * gridDim.x < p.n_tiles; it0 += B) prologue_pass(p, it0, B);
    return;
  }

  // loop state that survives a refresh call between two tiles (in-order claiming, see pipe_refresh_own)
  int it = 0;
  int tile = blockIdx.x;
  int refresh_prev = -1;  // due-list tile of the previous iteration, refreshed at the end of this one
  int refresh_now = -1;
  int cmd_next = 0;
restart:
  {
  // ---------------- per-thread constants of the tile loop ------------------------------------
  const int N = p.N, G = p.G;
  const int C = kC > 0 ? kC : p.C;
  const int half = C >> 1;
  const int ns = N + C;
  const int GN = G * N;                 // houses of a full tile
  const int T = p.hmax;                 // house threads of the CTA (multiple of 32)
  const int le = tid < GN ? (N == 1 ? tid : (int)__umulhi((unsigned)tid, p.div_magic)) : 0;  // tid / N
  const int li = tid - le * N;
  // own slot in message window 0 ([half halo | N houses | halo] per env); window 1 follows
  float4* const msg0 = reinterpret_cast<float4*>(smem_raw + p.off_msg) + (le * ns + half + li);
  const int msg_buf = G * ns;
  const bool halo_hi = li < C - half;   // my message is also the wrap-around halo after the last house
  const bool halo_lo = li >= N - half;  // ... and before the first one
  // warp-partial power sums: [2][G][part_stride]
  float* const part0 = reinterpret_cast<float*>(smem_raw + p.off_pw) + le * p.part_stride;
  const int part_buf = G * p.part_stride;
  float* const met0 = reinterpret_cast<float*>(smem_raw + p.off_met) + le * p.part_stride * 5;  // [2][G][part_stride][5]
  const int first_warp = (le * N) >> 5;
  const int my_part = warp - first_warp;
  const int nparts = ((le * N + N - 1) >> 5) - first_warp + 1;
  float* const row = reinterpret_cast<float*>(smem_raw + p.off_stage) + tid * p.F;  // rows contiguous like in HBM
  PipeEnv* const s_env = reinterpret_cast<PipeEnv*>(smem_raw + p.off_env);
  // cp.async input stage s (in_stride bytes each): [coef_a T x 16][coef_b T x 16][temps T x 8][coef_c T x 8][hvac T x 4]
  unsigned char* const in_a = smem_raw + p.off_in + tid * 16;
  unsigned char* const in_t = smem_raw + p.off_in + T * 32 + tid * 8;
  unsigned char* const in_h = smem_raw + p.off_in + T * 48 + tid * 4;
  const int in_stride = p.in_stride;
  const bool interp_mode = p.base_power_mode == MDR_BASE_INTERPOLATION;
  const int n_tiles = p.n_tiles;
  const int tile_stride = gridDim.x;
  const float inv_norm = p.f_inv_norm_reg_sig;

  auto tile_houses = [&](int tile) { return min(GN, (p.E - tile * G) * N); };
  auto issue_tile = [&](int tile, int s) {
    if (tid < tile_houses(tile)) {
      const unsigned h = (unsigned)tile * (unsigned)GN + (unsigned)tid;
      const int so = s * in_stride;
      cp_async_16(in_a + so, reinterpret_cast<const float4*>(p.coef_a) + h);
      cp_async_16(in_a + so + T * 16, reinterpret_cast<const float4*>(p.coef_b) + h);
      cp_async_8(in_t + so, reinterpret_cast<const float2*>(p.temps) + h);
      cp_async_8(in_t + so + T * 8, reinterpret_cast<const float2*>(p.coef_c) + h);
      cp_async_4(in_h + so, p.hvac + h);
    }
  };
  // the action byte of the next tile travels in a register (kept as loaded: converting here would
  // stall on the load instead of letting it fly)
  auto fetch_action = [&](int tile) -> int {
    if (kAct == MDR_ACT_ARRAY && tid < tile_houses(tile)) return p.actions[(unsigned)tile * (unsigned)GN + (unsigned)tid];
    return 0;
  };

  int any_due = 0;
  // strided lists: shared due queue (see DueQueue): a due tile is published one tile later, once its stores have landed
  DueQueue* const due_q = (interp_mode && !dyn) ? reinterpret_cast<DueQueue*>(p.workspace) : nullptr;
  int pend_tile = -1;
  if (it == 0 && refresh_now < 0) {  // (not when the loop is re-entered after a refresh)
    if (!dyn) asm volatile("griddepcontrol.wait;" ::: "memory");  // (house warps: after their loop-invariant set-up)
    MDR_CTA_STAMP(1);
    if (tile < n_tiles) {
      issue_tile(tile, 0);
      cmd_next = fetch_action(tile);
    }
    cp_async_commit();
    if (dyn) mbar_wait(&ctl.full[1], 0);  // position 1 of this CTA is known (behind the grid barrier)
  }
  refresh_now = -1;

  for (; tile < n_tiles; ++it, tile += tile_stride) {
    const int sbuf = it & 1;
    const int H = tile_houses(tile);
    const bool active = tid < H;
    const unsigned h = (unsigned)tile * (unsigned)GN + (unsigned)tid;
    const int e = tile * G + le;
    MDR_STAMP(0);
    int cmd = cmd_next;
    if (dyn) {
      // this warp has started tile it: position it + 3 may be claimed; position it + 1 has had two tiles to arrive
      if (lane == 0) mbar_arrive(&ctl.empty[(it + 3) & 3]);
      if (it > 0) mbar_wait(&ctl.full[(it + 1) & 3], ((it + 1) >> 2) & 1);
    }
    MDR_STAMP(8);
    const int next = dyn ? (ctl.ring_tile[(it + 1) & 3] & (kDuePhase - 1)) : tile + tile_stride;
    if (next < n_tiles) {
      issue_tile(next, sbuf ^ 1);
      cmd_next = fetch_action(next);
    }
    cp_async_commit();
    MDR_STAMP(9);
    // hand-over from the prologue warp (normally produced more than a tile ago).  Waiting here rather than
    // at the end of the tile (with the house threads prefetching od_temp themselves) measured the same.
    const int slot = it & ring_mask;
    const PipeEnv* const env_buf = s_env + slot * G;
    if (!dyn || it == 0) mbar_wait(&ctl.full[slot], (it >> ring_shift) & 1);  // (in-order claiming: waited for a tile ago)
    if (dyn && interp_mode && (ctl.ring_tile[slot] & kDuePhase) == 0 && ctl.tile_due[slot] != 0) {
      // a due tile met in address order: it is (or was) processed from the due list in this launch.  The iteration
      // keeps its shape -- one cp.async group, one barrier (the double-buffered windows count on it) -- and computes
      // nothing.
      cp_async_wait<1>();
      house_sync(T);
      refresh_now = refresh_prev;
      refresh_prev = -1;
      tile = next - tile_stride;  // (the loop header adds the stride)
      if (refresh_now >= 0) {
        ++it;
        tile += tile_stride;
        goto do_refresh;
      }
      continue;
    }
    const float od_old = env_buf[le].od_old;
    const float gain = env_buf[le].gain;  // 0 with solar gain off
    MDR_STAMP(1);
    cp_async_wait<1>();  // this thread's copies of the current tile have landed
    MDR_STAMP(2);

    // ---------------- phase A: per house ---------------------------------------------------
    float t_air = 0, t_mass = 0, target = 0, p_on = 0, deadband = 0, inv_lock = 1, pen = 0, pw = 0, terr = 0;
    int on = 0, lock = 0, sso = 0;
    float4* const msg = msg0 + sbuf * msg_buf;
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
      const float qa = (on ? cb.y : 0.0f) + gain;  // this step's solar gain (new datetime, :694)
      const float tss = od_old + qa * cb.x;
      const float x = tt.x - tss, y = tt.y - tss;
      t_air = tt.x + (ca4.x * x + ca4.y * y);
      t_mass = tt.y + (ca4.z * x + ca4.w * y);
      reinterpret_cast<float2*>(p.temps)[h] = make_float2(t_air, t_mass);
      p.hvac[h] = (sso << 2) | (lock << 1) | on;
      pw = on ? p_on : 0.0f;
      // SingleHouse.message :624-662 normalised as utils.py:842-868 (sso is scaled by the receiver)
      const float4 m = make_float4((t_air - target) * 0.2f, (float)sso, pw * inv_norm, p_on * inv_norm);
      msg[0] = m;
      if (halo_hi) msg[N] = m;
      if (halo_lo) msg[-N] = m;
      // utils.deadbandL2, utils.py:1266-1274
      const float hi = target + deadband * 0.5f, lo = target - deadband * 0.5f;
      if (hi < t_air) pen = (t_air - hi) * (t_air - hi);
      else if (lo > t_air) pen = (lo - t_air) * (lo - t_air);
      terr = t_air - target;
    }
    float* const part = part0 + sbuf * part_buf;
    {
      const int key = active ? le : -1;
      // fp32 partial sums: exact for integer-valued watts (every shipped capacity / COP; at most 224 houses < 2^24 W),
      // within fp32 rounding otherwise
      const float psum = segmented_sum<float>(pw, key, lane);
      const int prev_key = __shfl_up_sync(0xffffffffu, key, 1);
      const bool head = active && (lane == 0 || prev_key != key);
      if (head) part[my_part] = psum;
      if (kMetrics) {
        // per-house terms of main-deploy.py:124-139 / metrics.py:22-25; the reward sum follows from the penalty sum
        // (r_k = -(k_T pen_k + k_S dn^2), the second term is the same for every house of the env)
        const float s0 = segmented_sum<float>(pen, key, lane), s1 = segmented_sum<float>(terr, key, lane);
        const float s2 = segmented_sum<float>(fabsf(terr), key, lane), s3 = segmented_sum<float>(terr * terr, key, lane);
        float s4 = fabsf(terr);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float tv = __shfl_down_sync(0xffffffffu, s4, o);
          const int tk = __shfl_down_sync(0xffffffffu, key, o);
          if (lane + o < 32 && tk == key) s4 = fmaxf(s4, tv);
        }
        if (head) {
          float* d = met0 + (sbuf * part_buf + my_part) * 5;
          d[0] = s0; d[1] = s1; d[2] = s2; d[3] = s3; d[4] = s4;
        }
      }
    }
    MDR_STAMP(3);
    // the staging rows of this warp may still be read by the previous tile's bulk store (waited for
    // BEFORE the barrier: the two waits then overlap; after it they add up -- measured +5% on c4)
    if (kObs && it > 0) {
      if (lane == 0) bulk_wait_read_all();
      __syncwarp();
    }
    MDR_STAMP(4);
    // the only CTA-wide rendezvous of a tile: message window + power partials are complete.
    // (window / partials are double buffered, so nobody can overwrite what a slower warp still reads)
    house_sync(T);
    MDR_STAMP(5);

    float P = 0;
    if (active) {
      // an env of <= 224 houses spans at most 8 warps; same summation order as a counted loop
#pragma unroll
      for (int w = 0; w < 8; ++w)
        if (w < nparts) P += part[w];
      if (li == 0) p.cluster_power[e] = (double)P;
    }
    if (kMetrics && active && li == 0) {
      double t[4] = {0.0, 0.0, 0.0, 0.0};
      float mxf = 0.0f;
      const float* d = met0 + sbuf * part_buf * 5;
#pragma unroll
      for (int w = 0; w < 8; ++w)
        if (w < nparts) {
          t[0] += (double)d[w * 5 + 0]; t[1] += (double)d[w * 5 + 1]; t[2] += (double)d[w * 5 + 2]; t[3] += (double)d[w * 5 + 3];
          mxf = fmaxf(mxf, d[w * 5 + 4]);
        }
      const PipeEnv& pe = env_buf[le];
      const double mx = (double)mxf, Pd = (double)P;
      const double dn = (Pd - pe.s_old) * p.inv_n;
      double* m = p.metrics + (size_t)e * MDR_N_METRICS;
      atomicAdd(m + MDR_M_STEPS, 1.0);
      atomicAdd(m + MDR_M_SUM_MEAN_REWARD, -(t[0] * p.inv_n * p.k_temp + dn * dn * p.k_sig));
      atomicAdd(m + MDR_M_SUM_MEAN_TEMP_OFFSET, t[1] * p.inv_n);
      atomicAdd(m + MDR_M_SUM_MEAN_TEMP_ERROR, t[2] * p.inv_n);
      atomicAdd(m + MDR_M_SUM_SQ_TEMP_ERROR, t[3]);
      atomicAdd(m + MDR_M_SUM_SQ_MAX_TEMP_ERROR, mx * mx);
      // (mx >= 0: non-negative doubles order like their bit patterns)
      atomicMax(reinterpret_cast<unsigned long long*>(m + MDR_M_MAX_TEMP_ERROR), (unsigned long long)__double_as_longlong(mx));
      atomicAdd(m + MDR_M_SUM_OD_TEMP, pe.od_new);
      atomicAdd(m + MDR_M_SUM_CONSUMPTION, Pd);
      if (!pe.due) metrics_signal_terms(m, pe.sig_new, Pd);  // a due env: after its refresh (pipe_refresh_pass)
    }
    if (kObs && active) {
      // fast-path row: [T_air, T_mass, target, deadband, cap, on, lockout, sso, 1, signal, power | C x 4 messages]
      row[0] = (t_air - 20.0f) * 0.2f;
      row[1] = (t_mass - 20.0f) * 0.2f;
      row[2] = (target - 20.0f) * 0.2f;
      row[3] = deadband;
      row[4] = p_on * p.f_cop_over_def_cap;
      row[5] = (float)on;
      row[6] = (float)lock;
      row[7] = (float)sso * inv_lock;
      row[8] = 1.0f;
      row[10] = (float)((double)P * p.inv_norm_sig_agents);  // same bits as the generic kernel and the compact record
      // neighbours (:816-828) = the C window entries around this house, skipping itself
      const float4* win = msg - half;
      float* mrow = row + 11;
      uint4 dr = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < (kC > 0 ? kC : C); ++k) {
        const float4 m = win[k + (k >= half ? 1 : 0)];
        float kf = 1.0f;
        if (kDrops) {  // np.random.rand() > comm_defect_prob (:992): replayed, or one Philox block per four messages
          bool keep;
          if (p.msg_keep != nullptr) keep = p.msg_keep[(size_t)h * C + k] != 0;
          else {
            if ((k & 3) == 0) dr = drop_block(p, h, k >> 2);
            keep = drop_keep(p, dr, k);
          }
          kf = keep ? 1.0f : 0.0f;
        }
        mrow[4 * k + 0] = m.x * kf;
        mrow[4 * k + 1] = m.y * inv_lock * kf;
        mrow[4 * k + 2] = m.z * kf;
        mrow[4 * k + 3] = m.w * kf;
      }
    }
    MDR_STAMP(6);
    const int due_now = (!dyn && interp_mode) ? ctl.tile_due[slot] : 0;
    if (due_now) {
      any_due = 1;
      if (due_q == nullptr && tid == 0) {
        const int k = ctl.due_n;
        if (k < kMaxDue) ctl.due_list[k] = tile;
        ctl.due_n = k + 1;
      }
    }
    if (active) {
      // reg_signal_penalty :244-247 with the OLD signal; weighting :364-372
      const float dn = (float)((double)P - env_buf[le].s_old) * p.f_inv_n;
      if (p.reward != nullptr) reinterpret_cast<float*>(p.reward)[h] = -(pen * p.f_k_temp + dn * dn * p.f_k_sig);
      if (kObs) row[9] = env_buf[le].f_sig;
    }
    if (kObs) {
      const int wrow0 = warp * 32;
      const int nrows_w = min(32, H - wrow0);
      if (nrows_w > 0) {
        const int F = p.F;
        float* dst = reinterpret_cast<float*>(p.obs) + (size_t)((unsigned)tile * (unsigned)GN + (unsigned)wrow0) * F;
        const float* src = reinterpret_cast<const float*>(smem_raw + p.off_stage) + wrow0 * F;
        const uint32_t bytes = (uint32_t)(nrows_w * F * sizeof(float));
        const bool bulk_ok = ((reinterpret_cast<uintptr_t>(dst) | bytes) & 15) == 0;
        if (bulk_ok) {
          // measured alternatives (tools/microbench): a coalesced st.global.v4 copy loop is ~4% slower,
          // an L2 evict_first hint on this store changes nothing
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
    if (due_q != nullptr) {
      if (pend_tile >= 0) publish_due_tile(due_q, pend_tile, &ctl.pub_count[(it & 1) ^ 1], p.house_warps);
      pend_tile = due_now ? tile : -1;
    }
    // this warp is done with ring slot `slot`: let the prologue warp reuse it for tile it+ring
    __syncwarp();
    if (!dyn && lane == 0) mbar_arrive(&ctl.empty[slot]);
    MDR_STAMP(7);
    if (dyn) {
      if (interp_mode) {  // (CTA-uniform) the due-list tile of the previous iteration: its stores have had a tile to land
        refresh_now = refresh_prev;
        refresh_prev = (ctl.ring_tile[slot] & kDuePhase) ? tile : -1;
      }
      tile = next - tile_stride;  // (the loop header adds the stride)
      if (refresh_now >= 0) {
        ++it;
        tile += tile_stride;
        goto do_refresh;
      }
    }
  }
  cp_async_wait<0>();
  if (kObs && lane == 0) bulk_wait_read_all();
  MDR_CTA_STAMP(2);
  if (dyn && refresh_prev >= 0) {  // (the CTA's last tile came from the due list)
    refresh_now = refresh_prev;
    refresh_prev = -1;
    goto do_refresh;
  }
  // (with a shared due queue every CTA enters: it may have nothing due itself and still take tiles from the others)
  if (!dyn && interp_mode && (any_due || due_q != nullptr)) pipe_refresh_pass(p, le, li, pend_tile);  // CTA-uniform
  }
  goto finish;
do_refresh:
  pipe_refresh_own(p, refresh_now);
  goto restart;
finish:
  if (dyn && tid == 0) {  // the last CTA out leaves the claim header zeroed for the next launch
    DynHdr* const dh = dyn_hdr(p);
    if (atomicAdd(&dh->exited, 1u) == gridDim.x - 1) {
      dh->next_tile = 0;
      dh->arrived = 0;
      dh->n_due = 0;
      dh->exited = 0;
      __threadfence();
    }
  }
  MDR_CTA_STAMP(3);
}


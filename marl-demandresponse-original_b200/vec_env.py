"""Tensor fast path: E independent clusters x N houses resident in HBM, stepped by one fused
CUDA kernel per time step through the C ABI (include/mdr_b200.h).

``reset_tensor() -> obs[E, N, F]`` and
``step_tensor(actions[E, N]) -> (obs[E, N, F], reward[E, N], power[E], signal[E])``
are the batched generalisation of ``MADemandResponseEnv.reset/step``
(env/MA_DemandResponse.py:135-210) with the observation already normalised like
``utils.normStateDict`` (utils.py:740-880).  All returned tensors are CUDA tensors owned by
the environment and overwritten by the next call; nothing synchronises the host.

PyTorch is used for device memory and streams only; every per-step computation happens in
``csrc/mdr_kernels.cu``.  There is no CPU fallback.
"""
import copy
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .config_flatten import FlatConfig
from .default_config import INTERP_KEYS

_PRECISIONS = {"fp32": (_lib.F32, torch.float32), "fp64": (_lib.F64, torch.float64),
               "f32": (_lib.F32, torch.float32), "f64": (_lib.F64, torch.float64)}
TABLE_SIZE = 4199040


def load_interp_table(flat: FlatConfig):
    """np.load of power_grid_prop.base_power_parameters.interpolation.path_datafile
    (monteCarlo/interpolation.py:40); the blob is not shipped with the reference."""
    path = flat.interp_paths["path_datafile"]
    if not os.path.isfile(path):
        raise FileNotFoundError(
            "interpolation table %r not found (the reference lists it in .MISSING_LARGE_BLOBS); pass "
            "interp_table=... or use base_power_mode='constant'" % path)
    return np.load(path)


class VecDemandResponseEnv:
    def __init__(self, config, population, *, precision="fp32", device=None, interp_table=None, seed=0,
                 action_source="array", comm_table=None, test=False, with_obs=True, l2_persist=None):
        if not torch.cuda.is_available():
            raise _lib.MdrError("a CUDA device is required: this environment has no CPU fallback")
        self.lib = _lib.load()
        self.flat = config if isinstance(config, FlatConfig) else FlatConfig(config, test=test)
        self.precision_name = precision
        self.precision, self.dtype = _PRECISIONS[precision]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.seed, self.action_source, self.with_obs = int(seed), action_source, bool(with_obs)
        # L2 residency of the per-house state (None = MDR_L2_PERSIST env var; default OFF: measured 2x slower on c4)
        # ("table": the window covers the interpolation table instead -- 17 MB that the refreshes gather from)
        env_l2 = os.environ.get("MDR_L2_PERSIST", "0")
        self.l2_persist = ("table" if env_l2 == "table" else env_l2 == "1") if l2_persist is None else l2_persist
        self._l2_window = (0, 0, 1.0)
        self._flags, self._max_ctas = 0, 0
        self.n_envs = int(len(np.atleast_1d(population["t_epoch"])))
        self.n_houses = self.flat.n_houses
        self.n_comm = self.flat.n_comm
        self.n_features = self.flat.obs_width()
        self.step_index = 0
        self._keep = []  # replay tensors referenced by the last launch
        self._comm = None
        self._per_env_table = False
        if comm_table is None:
            comm_table = self.flat.explicit_comm_table() if self.flat.comm_mode_name != "random_fixed" else None
        self._alloc()
        self.load_population(population)
        if comm_table is not None:
            self.set_comm_table(comm_table)
        self._table = None
        if self.flat.base_power_mode == _lib.BASE["interpolation"]:
            if interp_table is None:
                interp_table = load_interp_table(self.flat)
            self.set_interp_table(interp_table)
        self._build_structs()

    # ------------------------------------------------------------------ memory
    def _alloc(self):
        e, n, dev, r = self.n_envs, self.n_houses, self.device, self.dtype
        f64, i32 = torch.float64, torch.int32
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)
        self.raw = {k: z(e, n, dtype=f64) for k in ("ua", "cm", "ca", "hm", "cap", "target", "deadband")}
        self.lockout_dur = z(e, n, dtype=i32)
        # everything the step kernel re-reads per house lives in ONE arena (coefficients, state,
        # action staging), so a single access-policy window can keep it L2-resident across steps
        rb = 4 if r == torch.float32 else 8
        sizes = [("coef_a", 4 * rb), ("coef_b", 4 * rb), ("coef_c", 2 * rb), ("temps", 2 * rb), ("hvac", 4), ("actions", 1)]
        offs, o = {}, 0
        for name, per_house in sizes:
            offs[name] = o
            o += (e * n * per_house + 255) // 256 * 256
        self._arena = torch.zeros(o, dtype=torch.uint8, device=dev)

        def carve(name, shape, dtype):
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            return self._arena[offs[name]:offs[name] + nbytes].view(dtype).view(*shape)

        self.coef_a, self.coef_b = carve("coef_a", (e, n, 4), r), carve("coef_b", (e, n, 4), r)
        self.coef_c, self.temps = carve("coef_c", (e, n, 2), r), carve("temps", (e, n, 2), r)
        self.hvac = carve("hvac", (e, n), i32)
        self.interp_key = z(e, n, dtype=i32)
        self.env = {k: z(e, dtype=f64) for k in ("phase", "od_temp", "solar_gain", "artificial_ratio",
                                                "max_power", "base_power", "signal", "cluster_power", "perlin_seed")}
        self.t_epoch = z(e, dtype=torch.int64)
        self.time_since_interp = z(e, dtype=i32)
        self.actions = carve("actions", (e, n), torch.uint8)
        self.obs = z(e, n, self.n_features, dtype=r) if self.with_obs else None
        self.reward = z(e, n, dtype=r)
        self.metrics = None  # [E, N_METRICS] fp64 accumulators, allocated by enable_metrics()
        self._pinned = None

    def load_population(self, pop):
        """Uploads a population dict (see population.py) and resets the step counter."""
        e, n, dev = self.n_envs, self.n_houses, self.device

        def house(key, dtype):
            a = np.asarray(pop[key]).reshape(e, n)
            return torch.as_tensor(np.ascontiguousarray(a)).to(device=dev, dtype=dtype)

        def env(key, dtype):
            a = np.asarray(pop[key]).reshape(e)
            return torch.as_tensor(np.ascontiguousarray(a)).to(device=dev, dtype=dtype)

        for k in self.raw:
            self.raw[k].copy_(house(k, torch.float64))
        self.lockout_dur.copy_(house("lockout_dur", torch.int32))
        self.temps[..., 0].copy_(house("t_air", self.dtype))
        self.temps[..., 1].copy_(house("t_mass", self.dtype))
        sso, on, lock = house("sso", torch.int32), house("on", torch.int32), house("lockout", torch.int32)
        self.hvac.copy_((sso << 2) | ((lock != 0).int() << 1) | (on != 0).int())
        for k in self.env:
            if k in pop:
                self.env[k].copy_(env(k, torch.float64))
        self.t_epoch.copy_(env("t_epoch", torch.int64))
        self.time_since_interp.copy_(env("time_since_interp", torch.int32))
        self.step_index = 0
        self._precomputed = False

    def set_comm_table(self, table):
        """int32 [N, C] shared by all envs, or [E, N, C] per env (random_fixed / random_sample)."""
        arr = np.ascontiguousarray(np.asarray(table, dtype=np.int32))
        if arr.size and (int(arr.min()) < 0 or int(arr.max()) >= self.n_houses):
            raise ValueError("comm table holds house ids outside [0, %d)" % self.n_houses)
        t = torch.as_tensor(arr).to(self.device)
        if t.dim() == 3:
            if t.shape != (self.n_envs, self.n_houses, self.n_comm):
                raise ValueError("per-env comm table must be [E, N, C]")
            self._per_env_table = True
        elif t.shape != (self.n_houses, self.n_comm):
            raise ValueError("comm table must be [N, C] = [%d, %d], got %s" % (self.n_houses, self.n_comm, tuple(t.shape)))
        else:
            self._per_env_table = False
        self._comm = t.contiguous()
        if hasattr(self, "cfg"):
            self._build_structs()

    def set_interp_table(self, table):
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(table).reshape(-1)))
        expect = int(np.prod([len(self.flat.interp_grid[k]) for k in INTERP_KEYS]))
        if t.numel() != expect:
            raise ValueError("interpolation table has %d entries, the grid needs %d" % (t.numel(), expect))
        self._table = t.to(device=self.device, dtype=self.dtype).contiguous()

    # ------------------------------------------------------------------ C structs
    def _build_structs(self):
        self.cfg = self.flat.to_struct(self.n_envs, self.precision, self.device.index, self.seed, self.action_source,
                                       per_env_table=self._per_env_table)
        if self.l2_persist:
            self._setup_l2_window()
        self.cfg.l2_window_base = C.c_void_p(self._l2_window[0]) if self._l2_window[1] else None
        self.cfg.l2_window_bytes, self.cfg.l2_hit_ratio = self._l2_window[1], self._l2_window[2]
        self.cfg.flags, self.cfg.max_ctas = self._flags, self._max_ctas
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        h = _lib.MdrHouses()
        for k in ("ua", "cm", "ca", "hm", "cap", "target", "deadband"):
            setattr(h, k, p(self.raw[k]))
        h.lockout_dur, h.coef_a, h.coef_b, h.coef_c = p(self.lockout_dur), p(self.coef_a), p(self.coef_b), p(self.coef_c)
        h.interp_key, h.temps, h.hvac = p(self.interp_key), p(self.temps), p(self.hvac)
        self.houses_s = h
        e = _lib.MdrEnvs()
        e.t_epoch, e.time_since_interp = p(self.t_epoch), p(self.time_since_interp)
        for k in self.env:
            setattr(e, k, p(self.env[k]))
        e.metrics = p(self.metrics)
        need = C.c_size_t()
        _lib.check(self.lib.mdr_workspace_bytes(C.byref(self.cfg), C.byref(need)), "mdr_workspace_bytes")
        if need.value and (getattr(self, "_workspace", None) is None or self._workspace.numel() < need.value):
            self._workspace = torch.zeros(need.value, dtype=torch.uint8, device=self.device)
        e.workspace = p(getattr(self, "_workspace", None))
        self.envs_s = e
        self.in_s = _lib.MdrStepInputs()
        self.out_s = _lib.MdrOutputs()
        self.out_s.obs, self.out_s.reward = p(self.obs), p(self.reward)
        self._refs = (C.byref(self.cfg), C.byref(self.houses_s), C.byref(self.envs_s), C.byref(self.in_s),
                      C.byref(self.out_s))

    def set_launch_options(self, *, no_pipeline=None, no_fused=None, no_pdl=None, no_cluster=None, static_tiles=None,
                           max_ctas=None):
        """MdrConfig.flags / max_ctas: pin the kernel choice (tests compare the pipelined, generic and fused kernels
        on the same inputs), choose between the pipelined kernel's fixed strided tile lists and in-order tile claiming
        (default: claiming from 14 tiles per CTA), and cap the persistent grid (so that every CTA walks many tiles in
        a small test)."""
        for bit, v in ((_lib.FLAG_NO_PIPELINE, no_pipeline), (_lib.FLAG_NO_FUSED, no_fused), (_lib.FLAG_NO_PDL, no_pdl),
                       (_lib.FLAG_NO_CLUSTER, no_cluster), (_lib.FLAG_STATIC_TILES, static_tiles)):
            if v is not None:
                self._flags = (self._flags | bit) if v else (self._flags & ~bit)
        if max_ctas is not None:
            self._max_ctas = int(max_ctas)
        self.cfg.flags, self.cfg.max_ctas = self._flags, self._max_ctas
        return self

    def _setup_l2_window(self):
        """Persisting-L2 carve-out + access policy window over the state arena (B200: 126 MB L2)."""
        granted, max_window = C.c_size_t(), C.c_size_t()
        target = self._arena
        if self.l2_persist == "table":
            if getattr(self, "_table", None) is None:
                self._l2_window = (0, 0, 1.0)
                return
            target = self._table.view(torch.uint8)
        nbytes = target.numel()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mdr_l2_persist_limit(self.device.index, nbytes, C.byref(granted), C.byref(max_window)),
                       "mdr_l2_persist_limit")
        window = min(nbytes, max_window.value)
        if granted.value == 0 or window == 0:
            self._l2_window = (0, 0, 1.0)
            return
        self._l2_window = (target.data_ptr(), window, min(1.0, granted.value / window))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, x, dtype, shape=None):
        if x is None:
            return None
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.asarray(x)))
        t = t.to(device=self.device, dtype=dtype).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            t = t.reshape(shape)
        return t

    def _set_inputs(self, actions=None, od_noise=None, signal_noise=None, interp_ids=None, msg_keep=None, comm=None):
        e, n, c = self.n_envs, self.n_houses, self.n_comm
        if actions is not None:
            if isinstance(actions, torch.Tensor) and actions.dtype == torch.uint8 and actions.is_contiguous() \
                    and actions.device == self.device and actions.numel() == e * n:
                act = actions
            else:
                a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions))
                act = (a.to(self.device) != 0).to(torch.uint8).reshape(e, n).contiguous()
        else:
            act = self.actions
        od = self._dev(od_noise, torch.float64, (e,))
        sn = self._dev(signal_noise, torch.float64, (e,))
        ids = self._dev(interp_ids, torch.int32, (e, self.flat.interp_nb_agents))
        mk = self._dev(msg_keep, torch.uint8, (e, n, c))
        cm = self._dev(comm, torch.int32)
        if cm is None:
            cm = self._comm
        elif cm.dim() == 2 and self._per_env_table:
            cm = cm.expand(e, n, c).contiguous()
        self._keep = [act, od, sn, ids, mk, cm, self._table]
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        s = self.in_s
        s.actions, s.od_noise, s.signal_noise, s.interp_ids = p(act), p(od), p(sn), p(ids)
        s.msg_keep, s.comm_table, s.interp_table = p(mk), p(cm), p(self._table)
        s.step_index = self.step_index
        s.step_counter = None

    # ------------------------------------------------------------------ API
    def precompute(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mdr_precompute(self._refs[0], self._refs[1], self._stream()), "mdr_precompute")
        self._precomputed = True

    def reset_tensor(self, *, signal_noise=None, interp_ids=None, msg_keep=None, comm=None):
        """Initial grid signal + initial observation for the loaded population
        (MADemandResponseEnv.build_environment tail :133 and reset :163-172)."""
        self.precompute()
        self._set_inputs(None, None, signal_noise, interp_ids, msg_keep, comm)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mdr_reset(*self._refs, self._stream()), "mdr_reset")
        return self.obs

    def reset_envs(self, mask=None, *, spec=None, draw_index=None):
        """Device-side (re)draw of the population and reset: `mask` [E] bool/uint8 selects the envs to
        reset (None = all); the others are not touched -- a rollout worker can restart finished episodes without a
        host round trip.  The draw follows utils.applyPropertyNoise / HVAC.__init__ / ClusterHouses.__init__ /
        PowerGrid.__init__ at the distribution level (Philox streams keyed by seed, draw_index, house/env).
        Returns the observation tensor (all envs), like reset_tensor(): the rows of the envs that were not reset are
        re-assembled from their unchanged state (identical bytes, except that with `comm_defect_prob > 0` or
        `random_sample` neighbours their message drops / neighbour sets are drawn anew, as at any observation)."""
        from .population import population_spec
        if spec is None:
            spec = population_spec(self.flat)
        self._draws = getattr(self, "_draws", 0) + 1
        di = self._draws if draw_index is None else int(draw_index)
        m = None
        if mask is not None:
            m = (mask if isinstance(mask, torch.Tensor) else torch.as_tensor(np.asarray(mask)))
            m = (m.to(self.device) != 0).to(torch.uint8).reshape(self.n_envs).contiguous()
        mp = C.c_void_p(m.data_ptr()) if m is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mdr_populate(self._refs[0], C.byref(spec), self._refs[1], self._refs[2], mp, di, self._stream()),
                       "mdr_populate")
        self.precompute()  # derived coefficients of every house (idempotent for the envs that were not re-drawn)
        self._set_inputs(None, None, None, None, None, None)
        self.in_s.env_mask = mp
        out_obs = self.out_s.obs
        self.out_s.obs = None
        try:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.mdr_reset(*self._refs, self._stream()), "mdr_reset")
        finally:
            self.in_s.env_mask = None
            self.out_s.obs = out_obs
        self._keep.append(m)
        return self.observe_tensor() if self.obs is not None else None

    def stagger_interp_clock(self, seed=0):
        """Spreads PowerGrid.time_since_last_interp (:1155, 1250-1255) of the envs uniformly over one refresh period
        (multiples of the time step), as it is in a rollout whose clusters were reset at different times: every step
        then refreshes ~dt/period of the envs instead of all of them every period/dt steps."""
        gen = torch.Generator(device=self.device).manual_seed(int(seed))
        slots = max(1, self.flat.interp_update_period // self.flat.time_step)
        k = torch.randint(0, slots, (self.n_envs,), generator=gen, device=self.device, dtype=torch.int32)
        self.time_since_interp.copy_(k * self.flat.time_step)
        return self.time_since_interp

    def observe_tensor(self, *, msg_keep=None, comm=None):
        """Observation of the current state; nothing advances."""
        if not self._precomputed:
            self.precompute()
        self._set_inputs(None, None, None, None, msg_keep, comm)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mdr_observe(*self._refs, self._stream()), "mdr_observe")
        return self.obs

    def step_tensor(self, actions=None, *, od_noise=None, signal_noise=None, interp_ids=None, msg_keep=None,
                    comm=None, n_steps=1, obs_out=None, reward_out=None, step_counter=None):
        """One env step for every cluster.  `actions` [E, N] (nonzero = ON) unless the env was
        built with an on-device action source.  Replay arguments feed host-drawn randomness
        (parity mode); when omitted the kernel draws from Philox / evaluates its own perlin."""
        if not self._precomputed:
            self.precompute()
        self._set_inputs(actions, od_noise, signal_noise, interp_ids, msg_keep, comm)
        # optional device-resident addend of the Philox step index (int64 [1]): CUDA-graph replays advance it on the device
        self.in_s.step_counter = C.c_void_p(step_counter.data_ptr()) if step_counter is not None else None
        obs, reward = self.obs, self.reward
        if obs_out is not None or reward_out is not None:
            # zero-copy: the kernel writes straight into the caller's rollout storage
            obs = self._check_out(obs_out, (self.n_envs, self.n_houses, self.n_features)) if obs_out is not None else obs
            reward = self._check_out(reward_out, (self.n_envs, self.n_houses), align=1) if reward_out is not None else reward
            self.out_s.obs = C.c_void_p(obs.data_ptr()) if obs is not None else None
            self.out_s.reward = C.c_void_p(reward.data_ptr())
        try:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.mdr_step(*self._refs, int(n_steps), self._stream()), "mdr_step")
        finally:
            if obs_out is not None or reward_out is not None:
                self.out_s.obs = C.c_void_p(self.obs.data_ptr()) if self.obs is not None else None
                self.out_s.reward = C.c_void_p(self.reward.data_ptr())
        self.step_index += int(n_steps)
        return obs, reward, self.env["cluster_power"], self.env["signal"]

    def _check_out(self, t, shape, align=16):
        # (the observation is written with 16-byte bulk stores; rewards with element stores)
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == self.device and t.dtype == self.dtype
                and t.is_contiguous() and tuple(t.shape) == tuple(shape) and t.data_ptr() % align == 0):
            raise ValueError("output tensor must be a contiguous, 16-byte aligned %s CUDA tensor of shape %s on %s"
                             % (self.dtype, tuple(shape), self.device))
        return t

    def run(self, n_steps):
        """`n_steps` steps with the on-device action source: the deploy loop of main-deploy.py:102-209.
        Configurations that need nothing from the host between steps (constant base power, individual_L2
        penalty, with_obs=False) run as ONE fused launch with the house state in registers; anything else
        is one launch per step, issued from C.  Returns (obs, reward, power, signal) of the LAST step."""
        if self.action_source == "array":
            raise ValueError("run() needs action_source='bangbang' or 'random'")
        return self.step_tensor(None, n_steps=n_steps)

    def enable_metrics(self, reset=True):
        """Allocates (or clears) the per-env accumulators of main-deploy.py:124-209 / metrics.py:22-47; they are
        updated on the device by run() / step_tensor() whenever the fused path applies (see mdr_step)."""
        if self.metrics is None:
            self.metrics = torch.zeros(self.n_envs, _lib.N_METRICS, dtype=torch.float64, device=self.device)
            self._build_structs()
        elif reset:
            self.metrics.zero_()
        return self.metrics

    def disable_metrics(self):
        self.metrics = None
        self._build_structs()

    def metrics_summary(self, metrics=None):
        """The figures main-deploy.py:176-209 prints, per env, from the accumulators ([E, N_METRICS] tensor,
        default: this env's own).  Returns a dict of [E] fp64 tensors."""
        m = self.metrics if metrics is None else metrics
        if m is None:
            raise ValueError("metrics are not enabled")
        col = {k: m[:, i] for i, k in enumerate(_lib.METRIC_NAMES)}
        steps = col["steps"].clamp(min=1.0)
        n = float(self.n_houses)
        return {
            "steps": col["steps"],
            "mean_reward": col["sum_mean_reward"] / steps,
            "mean_temp_offset": col["sum_mean_temp_offset"] / steps,
            "mean_temp_error": col["sum_mean_temp_error"] / steps,
            "max_temp_error": col["max_temp_error"],
            "rmse_temp": torch.sqrt(col["sum_sq_temp_error"] / (steps * n)),
            "rms_max_error_temp": torch.sqrt(col["sum_sq_max_temp_error"] / steps),
            "mean_od_temp": col["sum_od_temp"] / steps,
            "mean_signal": col["sum_signal"] / steps,
            "mean_consumption": col["sum_consumption"] / steps,
            "mean_signal_offset": col["sum_signal_offset"] / steps,
            "mean_signal_error": col["sum_signal_error"] / steps,
            "rmse_signal_per_agent": torch.sqrt(col["sum_sq_signal_error"] / steps) / n,
        }

    def host_pipeline(self, enable=True, n_threads=0, n_slices=0):
        """Creates (or drops) the MdrHostCtx of the pipelined host-buffer path: env-axis slices over two streams and,
        with the default observation layout, a compact 16-real record per house over PCIe expanded by `n_threads`
        host threads (0 = this process's CPU affinity count, at most 24).  Pin the process to the GPU's NUMA node first."""
        if getattr(self, "_host_ctx", None):
            _lib.check(self.lib.mdr_host_ctx_destroy(self._host_ctx), "mdr_host_ctx_destroy")
            self._host_ctx = None
        if enable:
            ctx = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(self.lib.mdr_host_ctx_create(C.byref(self.cfg), int(n_threads), int(n_slices), C.byref(ctx)),
                           "mdr_host_ctx_create")
            self._host_ctx = ctx
        return self

    def host_pipeline_info(self):
        if not getattr(self, "_host_ctx", None):
            return None
        t, s, b = C.c_int32(), C.c_int32(), C.c_size_t()
        _lib.check(self.lib.mdr_host_ctx_info(self._host_ctx, C.byref(t), C.byref(s), C.byref(b)), "mdr_host_ctx_info")
        return dict(threads=t.value, slices=s.value, compact_bytes=b.value)

    def host_transfer_bytes(self):
        """Device-to-host bytes of one step_host() call (observation + reward + per-env power and signal)."""
        e, n = self.n_envs, self.n_houses
        rb = 4 if self.dtype == torch.float32 else 8
        f = self.flat
        compact = bool(getattr(self, "_host_ctx", None)) and self.with_obs and f.comm_mode_name == "neighbours" \
            and f.state_flags == 0 and f.msg_flags == 0 and f.comm_defect_prob == 0 and n <= _lib.MAX_HOUSES_PER_CLUSTER
        per_house = (16 if compact else self.n_features) * rb if self.with_obs else 0
        return e * n * (per_house + rb) + 16 * e

    def __del__(self):
        try:
            if getattr(self, "_host_ctx", None):
                self.lib.mdr_host_ctx_destroy(self._host_ctx)
        except Exception:
            pass

    def step_host(self, host_actions, *, od_noise=None, signal_noise=None, interp_ids=None, msg_keep=None, comm=None,
                  want_obs=True):
        """End-to-end step with HOST buffers (what the dict API and the e2e benchmark call):
        H2D of the uint8 actions, the fused step, D2H of obs / reward / power / signal into pinned
        host memory, stream synchronised.  Returns numpy views of the pinned buffers.  With host_pipeline() enabled
        the transfers are sliced and overlapped and the observation travels as compact records (mdr_step_host)."""
        if not self._precomputed:
            self.precompute()
        e, n = self.n_envs, self.n_houses
        if self._pinned is None:
            pin = lambda *s, dtype: torch.empty(*s, dtype=dtype, pin_memory=True)
            self._pinned = dict(actions=pin(e, n, dtype=torch.uint8),
                                obs=pin(e, n, self.n_features, dtype=self.dtype) if self.with_obs else None,
                                reward=pin(e, n, dtype=self.dtype), power=pin(e, dtype=torch.float64),
                                signal=pin(e, dtype=torch.float64))
        pb = self._pinned
        ha = np.asarray(host_actions)
        if ha.dtype == np.bool_:
            ha = ha.view(np.uint8)
        # (the kernels read "nonzero = ON": uint8 / bool actions are copied as they are, one pass over 1 byte per house)
        np.copyto(pb["actions"].numpy(), ha.reshape(e, n) if ha.dtype == np.uint8 else (ha.reshape(e, n) != 0))
        self._set_inputs(None, od_noise, signal_noise, interp_ids, msg_keep, comm)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        skip_obs = not want_obs and self.obs is not None
        if skip_obs:  # the caller builds its observation from the state (dict API): no row assembly, no obs D2H
            self.out_s.obs = None
        try:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.mdr_step_host(*self._refs, p(pb["actions"]), None if skip_obs else p(pb["obs"]),
                                                  p(pb["reward"]), p(pb["power"]), p(pb["signal"]),
                                                  getattr(self, "_host_ctx", None), self._stream()),
                           "mdr_step_host")
        finally:
            if skip_obs:
                self.out_s.obs = C.c_void_p(self.obs.data_ptr())
        self.step_index += 1
        return (None if pb["obs"] is None or skip_obs else pb["obs"].numpy(), pb["reward"].numpy(), pb["power"].numpy(),
                pb["signal"].numpy())

    # ------------------------------------------------------------------ views of the state
    @property
    def t_air(self):
        return self.temps[..., 0]

    @property
    def t_mass(self):
        return self.temps[..., 1]

    @property
    def hvac_on(self):
        return self.hvac & 1

    @property
    def hvac_lockout(self):
        return (self.hvac >> 1) & 1

    @property
    def seconds_since_off(self):
        return self.hvac >> 2

    def launch_geometry(self):
        g, t, c, s, pl, cl = C.c_int32(), C.c_int32(), C.c_int32(), C.c_size_t(), C.c_int32(), C.c_int32()
        _lib.check(self.lib.mdr_launch_geometry(self._refs[0], int(self.with_obs), C.byref(g), C.byref(t), C.byref(c),
                                                C.byref(s), C.byref(pl), C.byref(cl)), "mdr_launch_geometry")
        return dict(envs_per_cta=g.value, threads=t.value, tiles=c.value, smem_bytes=s.value, cluster_size=cl.value,
                    kernel="mdr::step_wide_kernel (one CTA walks a whole env, no observation)" if pl.value == 2 else
                    ("mdr::step_pipe_split_kernel (persistent, software-pipelined, env split over a cluster)"
                            if pl.value and cl.value > 1 else "mdr::step_pipe_kernel (persistent, software-pipelined)") if pl.value else
                    ("mdr::big_update/env/finish_kernel (three launches)" if cl.value == 0 else "mdr::step_kernel"))

    # ------------------------------------------------------------------ checkpoint / copy
    _STATE = ("coef_a", "coef_b", "coef_c", "interp_key", "temps", "hvac", "lockout_dur", "t_epoch",
              "time_since_interp", "actions", "reward")

    def state_dict(self):
        sd = {k: getattr(self, k).clone() for k in self._STATE}
        sd.update({"raw." + k: v.clone() for k, v in self.raw.items()})
        sd.update({"env." + k: v.clone() for k, v in self.env.items()})
        if self.obs is not None:
            sd["obs"] = self.obs.clone()
        if self._comm is not None:
            sd["comm"] = self._comm.clone()
        if self.metrics is not None:
            sd["metrics"] = self.metrics.clone()
        sd["step_index"] = self.step_index
        sd["draws"] = getattr(self, "_draws", 0)  # population re-draws so far (reset_envs must not repeat a draw index)
        return sd

    def load_state_dict(self, sd):
        for k in self._STATE:
            getattr(self, k).copy_(sd[k])
        for k in self.raw:
            self.raw[k].copy_(sd["raw." + k])
        for k in self.env:
            self.env[k].copy_(sd["env." + k])
        if self.obs is not None and "obs" in sd:
            self.obs.copy_(sd["obs"])
        if "comm" in sd:
            self.set_comm_table(sd["comm"].cpu().numpy())
        if "metrics" in sd:
            self.enable_metrics()
            self.metrics.copy_(sd["metrics"])
        self.step_index = int(sd["step_index"])
        self._draws = int(sd.get("draws", getattr(self, "_draws", 0)))
        self._precomputed = True

    def __deepcopy__(self, memo):
        new = object.__new__(type(self))
        memo[id(self)] = new
        skip = {"lib", "cfg", "houses_s", "envs_s", "in_s", "out_s", "_refs", "_keep", "_pinned", "_arena", "_workspace",
                "_host_ctx", "coef_a",
                "coef_b", "coef_c", "temps", "hvac", "actions"}
        for k, v in self.__dict__.items():
            if k in skip:
                continue
            if isinstance(v, torch.Tensor):
                new.__dict__[k] = v.clone()
            elif isinstance(v, dict) and all(isinstance(x, torch.Tensor) for x in v.values()):
                new.__dict__[k] = {kk: vv.clone() for kk, vv in v.items()}
            elif k == "flat":
                new.__dict__[k] = v
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        new.lib, new._keep, new._pinned = self.lib, [], None
        # the arena-backed tensors are views: clone the arena and carve the same views out of the copy
        new._arena = self._arena.clone()
        base = self._arena.data_ptr()
        for k in ("coef_a", "coef_b", "coef_c", "temps", "hvac", "actions"):
            t = getattr(self, k)
            off = t.data_ptr() - base
            nbytes = t.numel() * t.element_size()
            new.__dict__[k] = new._arena[off:off + nbytes].view(t.dtype).view(*t.shape)
        new._build_structs()
        return new

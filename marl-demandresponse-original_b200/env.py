"""Drop-in ``MADemandResponseEnv``: the reference's dict-in / dict-out API
(env/MA_DemandResponse.py:37-390) on top of the CUDA step path.

``MADemandResponseEnv(config, test=False)``, ``reset() -> obs_dict`` and
``step(action_dict) -> (obs_dict, rewards_dict, dones_dict, info_dict)`` keep the reference's
names, argument meaning, dictionary keys, random-draw order (python ``random``) and
error behaviour, so the reference's agents and controllers run unchanged.  Host work per step:
marshal the action dict to a uint8 vector, draw the step's random numbers in the reference's
order, one ``mdr_step_host`` call, and build the python dicts from the returned state.
"""
import copy
import datetime as _dt
import random
import warnings

import numpy as np

from . import _lib
from .config_flatten import FlatConfig, from_epoch
from .perlin import Perlin
from .population import reference_order_population
from .vec_env import VecDemandResponseEnv


class _HVACView:
    def __init__(self, env, i):
        self._env, self.id = env, i

    COP = property(lambda s: s._env.flat.hvac_cop)
    latent_cooling_fraction = property(lambda s: s._env.flat.hvac_latent)
    cooling_capacity = property(lambda s: float(s._env._host["cap"][s.id]))
    lockout_duration = property(lambda s: int(s._env._host["lockout_dur"][s.id]))
    turned_on = property(lambda s: bool(s._env._host["on"][s.id]))
    lockout = property(lambda s: bool(s._env._host["lockout"][s.id]))
    seconds_since_off = property(lambda s: int(s._env._host["sso"][s.id]))
    max_consumption = property(lambda s: s.cooling_capacity / s.COP)

    def power_consumption(self):
        return self.max_consumption if self.turned_on else 0

    def get_Q(self):
        return -1 * self.cooling_capacity / (1 + self.latent_cooling_fraction) if self.turned_on else 0


class _HouseView:
    """Read-only stand-in for SingleHouse (callers only read attributes / use the dict keys)."""

    def __init__(self, env, i):
        self._env, self.id = env, i
        self.hvac = _HVACView(env, i)

    current_temp = property(lambda s: float(s._env._host["t_air"][s.id]))
    current_mass_temp = property(lambda s: float(s._env._host["t_mass"][s.id]))
    target_temp = property(lambda s: float(s._env._host["target"][s.id]))
    deadband = property(lambda s: float(s._env._host["deadband"][s.id]))
    Ua = property(lambda s: float(s._env._host["ua"][s.id]))
    Cm = property(lambda s: float(s._env._host["cm"][s.id]))
    Ca = property(lambda s: float(s._env._host["ca"][s.id]))
    Hm = property(lambda s: float(s._env._host["hm"][s.id]))
    current_solar_gain = property(lambda s: float(s._env._host["solar_gain"]))


class _ClusterView:
    def __init__(self, env):
        self._env = env
        self.houses = {i: _HouseView(env, i) for i in env.agent_ids}
        self.agent_ids = env.agent_ids
        self.nb_agents = env.nb_agents

    current_OD_temp = property(lambda s: float(s._env._host["od_temp"]))
    cluster_hvac_power = property(lambda s: float(s._env._host["cluster_power"]))
    max_power = property(lambda s: float(s._env._host["max_power"]))
    phase = property(lambda s: float(s._env._host["phase"]))

    @property
    def agent_communicators(self):
        t = self._env._comm_host
        return {i: [int(j) for j in t[i]] for i in self.agent_ids} if t is not None else {}


class _PowerGridView:
    def __init__(self, env):
        self._env = env
        self.cumulated_abs_noise = 0
        self.nb_steps = 0

    current_signal = property(lambda s: float(s._env._host["signal"]))
    base_power = property(lambda s: float(s._env._host["base_power"]))
    artificial_ratio = property(lambda s: float(s._env._host["artificial_ratio"]))
    max_power = property(lambda s: float(s._env._host["max_power"]))
    time_since_last_interp = property(lambda s: int(s._env._tsi))


class MADemandResponseEnv:
    """Multi agent demand response environment (B200 step path, reference API)."""

    def __init__(self, config, test=False, precision="fp64", device=None, interp_table=None):
        self.test = test
        self.config = config
        self.default_env_prop = config["default_env_prop"]
        self.default_house_prop = config["default_house_prop"]
        self.default_hvac_prop = config["default_hvac_prop"]
        if test:
            self.noise_house_prop = config["noise_house_prop_test"]
            self.noise_hvac_prop = config["noise_hvac_prop_test"]
        else:
            self.noise_house_prop = config["noise_house_prop"]
            self.noise_hvac_prop = config["noise_hvac_prop"]
        self._precision, self._device, self._interp_table = precision, device, interp_table
        self._vec = None
        self.build_environment()

    # ------------------------------------------------------------------
    def build_environment(self):
        """applyPropertyNoise + ClusterHouses + PowerGrid + the initial grid step (:98-133)."""
        self.flat = flat = FlatConfig(self.config, test=self.test)
        pop, table = reference_order_population(flat, random)
        self.start_datetime = from_epoch(pop["t_epoch"][0])
        self.datetime = self.start_datetime
        self.time_step = _dt.timedelta(seconds=flat.time_step)
        self.agent_ids = list(range(flat.n_houses))
        self.nb_agents = len(self.agent_ids)
        self.env_properties = copy.deepcopy(self.default_env_prop)
        self.env_properties.update(agent_ids=self.agent_ids, nb_hvac=self.nb_agents, start_datetime=self.start_datetime)
        self.env_properties["power_grid_prop"]["max_power"] = float(pop["max_power"][0])
        self._comm_host = table
        if table is None and flat.comm_mode_name == "neighbours":
            from .config_flatten import comm_table
            self._comm_host = comm_table("neighbours", flat.n_houses, flat.nb_agents_comm)
        if self._vec is None or self._vec.n_houses != flat.n_houses or self._vec.n_features != flat.obs_width():
            self._vec = VecDemandResponseEnv(flat, pop, precision=self._precision, device=self._device,
                                             interp_table=self._interp_table, comm_table=table)
        else:
            self._vec.flat = flat
            self._vec.load_population(pop)
            if table is not None:
                self._vec.set_comm_table(table)
            self._vec._build_structs()
        self._perlin = None
        if flat.signal_mode == _lib.SIG_PERLIN:
            sp = flat.signal_params
            self._perlin = Perlin(1, sp["nb_octaves"], sp["octaves_step"], sp["period"], float(pop["perlin_seed"][0]))
        self._tsi = flat.interp_update_period + 1
        self._static = {k: np.asarray(pop[k][0]) for k in ("ua", "cm", "ca", "hm", "cap", "target", "deadband",
                                                            "lockout_dur")}
        self._static_lists = {k: (v.astype(np.int64) if k == "lockout_dur" else v.astype(np.float64)).tolist()
                              for k, v in self._static.items()}
        self._static.update(max_power=float(pop["max_power"][0]), phase=float(pop["phase"][0]),
                            artificial_ratio=float(pop["artificial_ratio"][0]))
        # PowerGrid.step(start_datetime) at :133 -- interpolation ids are drawn first (:1214), then perlin
        ids, noise = self._grid_draws(self.datetime)
        placeholder = None
        if flat.comm_mode_name == "random_sample":  # neighbour sets are only drawn when an observation is made
            placeholder = np.zeros((1, flat.n_houses, flat.n_comm), np.int32)
        self._vec.reset_tensor(signal_noise=noise, interp_ids=ids, comm=placeholder)
        self._pull_state()
        self.cluster = _ClusterView(self)
        self.power_grid = _PowerGridView(self)

    def _grid_draws(self, date_time):
        flat = self.flat
        ids = None
        if flat.base_power_mode == _lib.BASE["interpolation"]:
            self._tsi += flat.time_step
            if self._tsi >= flat.interp_update_period:
                self._tsi = 0
                if flat.n_houses > flat.interp_nb_agents:
                    ids = np.asarray(random.choices(self.agent_ids, k=flat.interp_nb_agents), dtype=np.int32)[None]
        noise = None
        if self._perlin is not None:
            unix = (date_time - _dt.datetime(1970, 1, 1)).total_seconds() % 86400  # time.mktime(...) % 86400 with TZ=UTC
            noise = np.array([self._perlin.calculate_noise(unix)])
        return ids, noise

    def _message_draws(self):
        """random_sample neighbour sets (:976-983) and message-drop uniforms (:992), house by house."""
        flat = self.flat
        comm, keep = None, None
        n, c = flat.n_houses, flat.n_comm
        if flat.comm_mode_name == "random_sample" or flat.comm_defect_prob > 0:
            comm_rows, keep = [], np.ones((1, n, c), np.uint8)
            for i in self.agent_ids:
                if flat.comm_mode_name == "random_sample":
                    possible = [j for j in self.agent_ids if j != i]
                    comm_rows.append(random.sample(possible, k=c))
                for k in range(c):
                    keep[0, i, k] = np.random.rand() > flat.comm_defect_prob
            if comm_rows:
                comm = np.asarray(comm_rows, dtype=np.int32)[None]
                self._comm_host = comm[0]
        elif c > 0:
            # the reference calls np.random.rand() once per message at every observation (:992) even when nothing can be
            # dropped: consume the same N*C uniforms so that learners sharing the global numpy stream
            # (agents/network.py:161,172, agents/ddpg.py:268) see the reference's sequence under the same seed
            np.random.rand(n * c)
        return comm, keep

    _ENV_SCALARS = ("od_temp", "signal", "cluster_power", "base_power", "solar_gain")

    def _pull_state(self):
        """Host copy of the state the dict API exposes: three asynchronous device-to-host copies into pinned memory
        (temperatures, packed HVAC state, the five per-env scalars gathered into one tensor) and one synchronisation."""
        import torch
        v = self._vec
        if getattr(self, "_pin", None) is None or self._pin[0].shape[0] != v.n_houses:
            self._pin = (torch.empty(v.n_houses, 2, dtype=v.dtype, pin_memory=True),
                         torch.empty(v.n_houses, dtype=torch.int32, pin_memory=True),
                         torch.empty(len(self._ENV_SCALARS), dtype=torch.float64, pin_memory=True))
        pt, ph, ps = self._pin
        pt.copy_(v.temps[0], non_blocking=True)
        ph.copy_(v.hvac[0], non_blocking=True)
        ps.copy_(torch.stack([v.env[k][0] for k in self._ENV_SCALARS]), non_blocking=True)
        torch.cuda.current_stream(v.device).synchronize()  # one synchronisation for the three copies
        temps = pt.numpy().astype(np.float64)
        hv = ph.numpy().copy()
        scal = ps.tolist()
        h = dict(self._static)
        h.update(t_air=temps[:, 0], t_mass=temps[:, 1], on=hv & 1, lockout=(hv >> 1) & 1, sso=hv >> 2)
        h.update(zip(self._ENV_SCALARS, scal))
        self._host = h

    # ------------------------------------------------------------------
    def reset(self):
        """Reset the environment; returns obs_dict (:135-172)."""
        self.build_environment()
        comm, keep = self._message_draws()
        self._msg_keep, self._msg_comm = keep, comm
        return self._make_obs_dict()

    def step(self, action_dict):
        """Take a step in time for each TCL, given actions of TCL agents (:174-210)."""
        flat = self.flat
        self.datetime += self.time_step
        actions = np.zeros(flat.n_houses, np.uint8)
        for i in self.agent_ids:
            if i in action_dict.keys():
                actions[i] = 1 if action_dict[i] else 0
            else:
                warnings.warn("HVAC in house {} did not receive any command.".format(i))
        od_noise = np.array([random.gauss(0, flat.temp_std)])      # compute_OD_temp :1079
        comm, keep = self._message_draws()                          # make_cluster_obs_dict :976-1002
        ids, noise = self._grid_draws(self.datetime)                # PowerGrid.step :1236-1316
        self._msg_keep, self._msg_comm = keep, comm
        _, reward, power, _ = self._vec.step_host(actions[None], od_noise=od_noise, signal_noise=noise,
                                                  interp_ids=ids, msg_keep=keep, comm=comm, want_obs=False)
        self._pull_state()
        if self._perlin is not None:
            self.power_grid.nb_steps += 1
        obs_dict = self._make_obs_dict()
        rewards_dict = dict(zip(self.agent_ids, reward[0].astype(np.float64).tolist()))
        dones_dict = dict.fromkeys(self.agent_ids, False)
        info_dict = {"cluster_hvac_power": float(power[0])}
        return obs_dict, rewards_dict, dones_dict, info_dict

    _OBS_KEYS = ("OD_temp", "datetime", "house_temp", "house_mass_temp", "hvac_turned_on", "hvac_seconds_since_off",
                 "hvac_lockout", "house_target_temp", "house_deadband", "house_Ua", "house_Cm", "house_Ca", "house_Hm",
                 "house_solar_gain", "hvac_COP", "hvac_cooling_capacity", "hvac_latent_cooling_fraction",
                 "hvac_lockout_duration", "message", "reg_signal", "cluster_hvac_power")

    def _make_obs_dict(self):
        """make_cluster_obs_dict + merge_cluster_powergrid_obs (:904-1003, :212-232) from host copies.  Same keys, key
        order and python types as the reference; built column-wise (`tolist` + `zip`) because at this point the python
        dictionaries, not the step, are what a drop-in user waits for.  A sender's message dict is built once and
        shared by its receivers (the reference builds equal copies; nothing downstream mutates them)."""
        from itertools import repeat
        h, flat = self._host, self.flat
        mp = self.default_env_prop["message_properties"]
        n, c = flat.n_houses, flat.n_comm
        p_max = (h["cap"] / flat.hvac_cop)
        t_air, target, sso, on = h["t_air"], h["target"], h["sso"], h["on"]
        # one message per sender (SingleHouse.message :624-662)
        cols = [("current_temp_diff_to_target", (t_air - target).tolist()), ("hvac_seconds_since_off", sso.tolist()),
                ("hvac_curr_consumption", [pm if o else 0 for pm, o in zip(p_max.tolist(), on.tolist())]),
                ("hvac_max_consumption", p_max.tolist()), ("hvac_lockout_duration", h["lockout_dur"].tolist())]
        if mp["thermal"]:
            cols += [("house_Ua", h["ua"].tolist()), ("house_Cm", h["cm"].tolist()), ("house_Ca", h["ca"].tolist()),
                     ("house_Hm", h["hm"].tolist())]
        if mp["hvac"]:
            cols += [("hvac_COP", [flat.hvac_cop] * n), ("hvac_cooling_capacity", h["cap"].tolist()),
                     ("hvac_latent_cooling_fraction", [flat.hvac_latent] * n)]
        mkeys = [k for k, _ in cols]
        sent = [dict(zip(mkeys, row)) for row in zip(*(v for _, v in cols))]
        messages = [[]] * n
        if c > 0:
            comm = self._comm_host
            comm_rows = comm.tolist() if hasattr(comm, "tolist") else [list(r) for r in comm]
            keep = self._msg_keep
            if keep is None:
                messages = [[sent[j] for j in row[:c]] for row in comm_rows]
            else:
                dropped = dict.fromkeys(mkeys, 0)  # np.random.rand() <= comm_defect_prob, :992-1001
                keep_rows = keep[0].tolist()
                messages = [[sent[j] if ok else dropped for j, ok in zip(row[:c], krow)] for row, krow in zip(comm_rows, keep_rows)]
        columns = (repeat(h["od_temp"]), repeat(self.datetime), t_air.tolist(), h["t_mass"].tolist(),
                   [bool(x) for x in on.tolist()], sso.tolist(), [bool(x) for x in h["lockout"].tolist()],
                   self._static_lists["target"], self._static_lists["deadband"], self._static_lists["ua"],
                   self._static_lists["cm"], self._static_lists["ca"], self._static_lists["hm"], repeat(h["solar_gain"]),
                   repeat(flat.hvac_cop), self._static_lists["cap"], repeat(flat.hvac_latent),
                   self._static_lists["lockout_dur"], messages, repeat(h["signal"]), repeat(h["cluster_power"]))
        keys = self._OBS_KEYS
        return {i: dict(zip(keys, row)) for i, row in zip(self.agent_ids, zip(*columns))}

    # tensor view of the last observation: the normStateDict matrix [N, F] on the device
    def obs_tensor(self):
        return self._vec.observe_tensor(msg_keep=self._msg_keep, comm=self._msg_comm)[0]

    def __deepcopy__(self, memo):
        new = object.__new__(type(self))
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("cluster", "power_grid", "_pin"):
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        new._pin = None  # pinned staging buffers are per object
        new.cluster = _ClusterView(new)
        new.power_grid = _PowerGridView(new)
        return new

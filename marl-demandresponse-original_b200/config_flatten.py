"""Flattens the reference's nested config dict (config.py; read by MADemandResponseEnv at
env/MA_DemandResponse.py:86-94) into the POD ``MdrConfig`` of the C ABI, once, at construction.
Raises the same ``ValueError`` s the reference constructors raise for unknown modes
(:249, :324, :898, :1169, :1306, :859-864)."""
import datetime as _dt
import math

import numpy as np

from . import _lib
from .default_config import INTERP_GRID, INTERP_KEYS

EPOCH = _dt.datetime(1970, 1, 1)


def comm_table(mode, n, nb_agents_comm, row_size=5, distance_comm=2, sampler=None):
    """Neighbour ids int32 [N, C] of ClusterHouses.build_agent_comm_links (:806-902).
    ``sampler(possible_ids, k)`` supplies random.sample for the random_fixed mode."""
    nb_comm = int(min(nb_agents_comm, n - 1))
    if mode == "neighbours":
        ids = np.arange(n)[:, None]
        before = np.arange(nb_comm // 2)[None, :] - nb_comm // 2
        after = np.arange(int(math.ceil(nb_comm / 2)))[None, :] + 1
        table = np.concatenate([(ids + before) % n, (ids + after) % n], axis=1)
    elif mode == "closed_groups":
        rows = []
        for i in range(n):
            base = i - (i % (nb_comm + 1))
            if base + nb_comm <= n:
                group = list(range(base, base + nb_agents_comm + 1))
            else:
                group = list(range(n - nb_comm - 1, n))
            group.remove(i)
            rows.append(group)
        table = np.asarray(rows).reshape(n, -1)
        if table.size and int(table.max()) >= n:
            # reference quirk (:834): `base + nb_comm <= nb_agents` lets the last full group run one id past the end
            # when base + nb_comm == nb_agents; the reference then dies with KeyError in make_cluster_obs_dict
            # (:958, self.houses[id]).  Fail the same way, but at construction instead of at the first observation.
            raise KeyError(int(table.max()))
    elif mode == "random_fixed":
        if sampler is None:
            raise ValueError("random_fixed needs a sampler")
        rows = []
        for i in range(n):
            possible = [j for j in range(n) if j != i]
            rows.append(list(sampler(possible, nb_comm)))
        table = np.asarray(rows).reshape(n, nb_comm)
    elif mode == "neighbours_2D":
        if n % row_size != 0:
            raise ValueError("Neighbours 2D row_size must be a divisor of nb_agents")
        max_y = n // row_size
        if distance_comm >= (row_size + 1) // 2 or distance_comm >= (max_y + 1) // 2:
            raise ValueError(
                "Neighbours 2D distance_comm ({}) must be strictly smaller than (row_size+1) / 2 ({}) and "
                "(max_y+1) / 2 ({})".format(distance_comm, (row_size + 1) // 2, (max_y + 1) // 2))
        offs = [(dx, dy) for dx in range(-distance_comm, distance_comm + 1)
                for dy in range(-distance_comm, distance_comm + 1)
                if abs(dx) + abs(dy) <= distance_comm and (dx, dy) != (0, 0)]
        x, y = np.arange(n) % row_size, np.arange(n) // row_size
        cols = [((y + dy) % max_y) * row_size + ((x + dx) % row_size) for dx, dy in offs]
        table = np.stack(cols, axis=1)
    elif mode in ("no_message", "random_sample"):
        table = np.zeros((n, 0 if mode == "no_message" else nb_comm), dtype=np.int32)
    else:
        raise ValueError("Cluster property: unknown agents_comm_mode '{}'.".format(mode))
    return np.ascontiguousarray(table, dtype=np.int32)


class FlatConfig:
    """Everything static about one environment configuration."""

    def __init__(self, config: dict, test: bool = False):
        self.config = config
        env_prop = config["default_env_prop"]
        self.house_def = config["default_house_prop"]
        self.hvac_def = config["default_hvac_prop"]
        self.noise_house = config["noise_house_prop_test" if test else "noise_house_prop"]
        self.noise_hvac = config["noise_hvac_prop_test" if test else "noise_hvac_prop"]
        self.env_prop = env_prop
        cp, gp, rp = env_prop["cluster_prop"], env_prop["power_grid_prop"], env_prop["reward_prop"]
        self.n_houses = int(cp["nb_agents"])
        self.time_step = int(env_prop["time_step"])
        self.comm_mode_name = cp["agents_comm_mode"]
        p2d = cp["agents_comm_parameters"]["neighbours_2D"]
        self.row_size, self.distance_comm = p2d["row_size"], p2d["distance_comm"]
        self.nb_agents_comm = int(cp["nb_agents_comm"])
        if self.comm_mode_name not in ("neighbours", "closed_groups", "random_sample", "random_fixed",
                                       "neighbours_2D", "no_message"):
            raise ValueError("Cluster property: unknown agents_comm_mode '{}'.".format(self.comm_mode_name))
        n = self.n_houses
        if self.comm_mode_name == "neighbours_2D":
            self.n_comm = comm_table("neighbours_2D", n, self.nb_agents_comm, self.row_size, self.distance_comm).shape[1]
        elif self.comm_mode_name == "no_message":
            self.n_comm = 0
        elif self.comm_mode_name == "closed_groups":
            self.n_comm = comm_table("closed_groups", n, self.nb_agents_comm).shape[1]
        else:
            self.n_comm = int(min(self.nb_agents_comm, n - 1))
        self.comm_defect_prob = float(cp["comm_defect_prob"])
        tm = cp["temp_parameters"][cp["temp_mode"]]
        self.day_temp, self.night_temp = float(tm["day_temp"]), float(tm["night_temp"])
        self.temp_std, self.random_phase_offset = float(tm["temp_std"]), bool(tm["random_phase_offset"])
        sp, mp = env_prop["state_properties"], env_prop["message_properties"]
        self.state_flags = ((_lib.STATE_HOUR if sp["hour"] else 0) | (_lib.STATE_DAY if sp["day"] else 0)
                            | (_lib.STATE_SOLAR if sp["solar_gain"] else 0)
                            | (_lib.STATE_THERMAL if sp["thermal"] else 0) | (_lib.STATE_HVAC if sp["hvac"] else 0))
        self.msg_flags = (_lib.MSG_THERMAL if mp["thermal"] else 0) | (_lib.MSG_HVAC if mp["hvac"] else 0)
        if rp["sig_penalty_mode"] != "common_L2":
            raise ValueError("Unknown signal penalty mode: {}".format(rp["sig_penalty_mode"]))
        if rp["temp_penalty_mode"] not in _lib.PEN:
            raise ValueError("Unknown temperature penalty mode: {}".format(rp["temp_penalty_mode"]))
        self.temp_penalty_mode = _lib.PEN[rp["temp_penalty_mode"]]
        mix = rp["temp_penalty_parameters"]["mixture"]
        self.mix = (float(mix["alpha_ind_L2"]), float(mix["alpha_common_L2"]), float(mix["alpha_common_max"]))
        self.alpha_temp, self.alpha_sig = float(rp["alpha_temp"]), float(rp["alpha_sig"])
        self.norm_reg_sig = float(rp["norm_reg_sig"])
        # compute_rewards normalisers, :346-356 (deadbandL2 with a zero deadband)
        self.norm_temp_penalty = float((self.house_def["target_temp"] + 1 - self.house_def["target_temp"]) ** 2)
        self.norm_sig_penalty = float((self.norm_reg_sig - 0.75 * self.norm_reg_sig) ** 2)
        self.solar_gain = bool(self.house_def["solar_gain_bool"])
        if gp["base_power_mode"] not in _lib.BASE:
            raise ValueError("The base_power_mode parameter in the config file can only be 'constant' or "
                             "'interpolation'. It is currently: {}".format(gp["base_power_mode"]))
        self.base_power_mode = _lib.BASE[gp["base_power_mode"]]
        self.avg_power_per_hvac = float(gp["base_power_parameters"]["constant"]["avg_power_per_hvac"])
        ip = gp["base_power_parameters"]["interpolation"]
        self.interp_update_period, self.interp_nb_agents = int(ip["interp_update_period"]), int(ip["interp_nb_agents"])
        self.interp_paths = ip
        self.signal_mode_name = gp["signal_mode"]
        self.signal_params = gp["signal_parameters"].get(self.signal_mode_name)
        if self.signal_mode_name == "flat":
            self.signal_mode = _lib.SIG_FLAT
        elif self.signal_mode_name == "sinusoidals":
            self.signal_mode = _lib.SIG_SINUSOIDALS
            if len(self.signal_params["periods"]) != len(self.signal_params["amplitude_ratios"]):
                raise ValueError("Power grid signal parameters: periods and amplitude_ratios lists should have the "
                                 "same length.")
        elif self.signal_mode_name == "regular_steps":
            self.signal_mode = _lib.SIG_REGULAR_STEPS
        elif "perlin" in self.signal_mode_name:
            self.signal_mode = _lib.SIG_PERLIN
        else:
            raise ValueError("Invalid power grid signal mode: {}. Change value in the config file.".format(
                self.signal_mode_name))
        if self.signal_params is None:
            raise KeyError(self.signal_mode_name)
        self.artificial_ratio = float(gp["artificial_ratio"])
        self.artificial_signal_ratio_range = float(gp["artificial_signal_ratio_range"])
        self.start_datetime = _dt.datetime.strptime(env_prop["start_datetime"], "%Y-%m-%d %H:%M:%S")
        self.start_datetime_mode = env_prop["start_datetime_mode"]
        if self.start_datetime_mode not in ("random", "fixed"):
            raise ValueError("start_datetime_mode in default_env_prop in config.py must be random or fixed.")
        self.hvac_cop = float(self.hvac_def["COP"])
        self.hvac_latent = float(self.hvac_def["latent_cooling_fraction"])
        if self.hvac_latent > 1 or self.hvac_latent < 0:
            raise ValueError("Latent cooling fraction must be between 0 and 1. Current value: {}.".format(self.hvac_latent))
        if self.hvac_cop < 0:
            raise ValueError("Coefficient of performance (COP) must be positive. Current value: {}.".format(self.hvac_cop))
        self.interp_grid = {k: list(map(float, INTERP_GRID[k])) for k in INTERP_KEYS}

    # ------------------------------------------------------------------------------
    def explicit_comm_table(self, sampler=None):
        """None for the implicit `neighbours` / `no_message` modes, else int32 [N, C]."""
        if self.comm_mode_name in ("neighbours", "no_message", "random_sample"):
            return None
        return comm_table(self.comm_mode_name, self.n_houses, self.nb_agents_comm, self.row_size, self.distance_comm,
                          sampler)

    def comm_mode_id(self, per_env_table=False):
        if self.comm_mode_name == "neighbours":
            return _lib.COMM_NEIGHBOURS
        if self.comm_mode_name == "no_message":
            return _lib.COMM_NONE
        if self.comm_mode_name == "random_sample" or per_env_table:
            return _lib.COMM_TABLE_PER_ENV
        return _lib.COMM_TABLE

    def obs_width(self):
        own = 11
        if self.state_flags & _lib.STATE_THERMAL:
            own += 5
        for flag, k in ((_lib.STATE_DAY, 2), (_lib.STATE_HOUR, 2), (_lib.STATE_SOLAR, 1), (_lib.STATE_HVAC, 2)):
            if self.state_flags & flag:
                own += k
        msg = 4 + (4 if self.msg_flags & _lib.MSG_THERMAL else 0) + (3 if self.msg_flags & _lib.MSG_HVAC else 0)
        return own + msg * self.n_comm

    def to_struct(self, n_envs, precision, device, seed=0, action_source="array", per_env_table=False):
        c = _lib.MdrConfig()
        c.abi_version, c.device, c.precision = _lib.MDR_ABI_VERSION, int(device), int(precision)
        c.n_envs, c.n_houses, c.n_comm, c.n_features = int(n_envs), self.n_houses, self.n_comm, self.obs_width()
        c.time_step, c.comm_mode = self.time_step, self.comm_mode_id(per_env_table)
        c.state_flags, c.msg_flags, c.temp_penalty_mode = self.state_flags, self.msg_flags, self.temp_penalty_mode
        c.solar_gain, c.base_power_mode, c.signal_mode = int(self.solar_gain), self.base_power_mode, self.signal_mode
        c.interp_update_period, c.interp_nb_agents = self.interp_update_period, self.interp_nb_agents
        c.action_source = _lib.ACT[action_source]
        c.obs_norm_agents = self.n_houses
        c.alpha_temp, c.alpha_sig = self.alpha_temp, self.alpha_sig
        c.norm_temp_penalty, c.norm_sig_penalty = self.norm_temp_penalty, self.norm_sig_penalty
        c.mix_alpha_ind, c.mix_alpha_common, c.mix_alpha_max = self.mix
        c.norm_reg_sig = self.norm_reg_sig
        c.def_ua, c.def_cm = float(self.house_def["Ua"]), float(self.house_def["Cm"])
        c.def_ca, c.def_hm = float(self.house_def["Ca"]), float(self.house_def["Hm"])
        c.def_cop, c.def_latent = self.hvac_cop, self.hvac_latent
        c.def_cap = float(self.hvac_def["cooling_capacity"])
        c.hvac_cop, c.hvac_latent = self.hvac_cop, self.hvac_latent
        c.day_temp, c.night_temp, c.temp_std = self.day_temp, self.night_temp, self.temp_std
        c.window_area, c.shading_coeff = float(self.house_def["window_area"]), float(self.house_def["shading_coeff"])
        c.avg_power_per_hvac = self.avg_power_per_hvac
        sp = self.signal_params
        if self.signal_mode == _lib.SIG_SINUSOIDALS:
            if len(sp["periods"]) > _lib.MAX_SINUSOIDS:
                raise ValueError("at most %d sinusoids are supported" % _lib.MAX_SINUSOIDS)
            c.n_sinusoids = len(sp["periods"])
            for i, (per, ratio) in enumerate(zip(sp["periods"], sp["amplitude_ratios"])):
                c.sin_periods[i], c.sin_ratios[i] = float(per), float(ratio)
        elif self.signal_mode == _lib.SIG_REGULAR_STEPS:
            c.steps_amplitude_per_hvac, c.steps_period = float(sp["amplitude_per_hvac"]), float(sp["period"])
        elif self.signal_mode == _lib.SIG_PERLIN:
            c.perlin_amplitude, c.perlin_period = float(sp["amplitude_ratios"]), float(sp["period"])
            c.perlin_nb_octaves, c.perlin_octaves_step = int(sp["nb_octaves"]), int(sp["octaves_step"])
        c.comm_defect_prob = self.comm_defect_prob
        for d, k in enumerate(INTERP_KEYS):
            axis = self.interp_grid[k]
            c.interp_dims[d] = len(axis)
            for j, v in enumerate(axis):
                c.interp_axes[d][j] = v
        c.seed = int(seed) & (2**64 - 1)
        return c


def epoch_seconds(d: _dt.datetime) -> int:
    return int((d - EPOCH).total_seconds())


def from_epoch(t: int) -> _dt.datetime:
    return EPOCH + _dt.timedelta(seconds=int(t))

"""Environment-relevant defaults of the reference's ``config.py`` (house :12-26, house noise
:27-128, HVAC :130-196, env :198-419), restated as data so that the package works on a box
where the reference tree is absent.  ``tests/test_host.py`` checks these dicts for equality
against the reference's ``config_dict`` whenever ``/root/reference`` is present.

Only the sub-dicts ``MADemandResponseEnv`` reads (env/MA_DemandResponse.py:86-94) are kept;
agent / training hyper-parameters are outside the step path.
"""
import copy


def _house_noise(modes):
    return {k: dict(std_start_temp=a, std_target_temp=b, factor_thermo_low=lo, factor_thermo_high=hi)
            for k, (a, b, lo, hi) in modes.items()}


def _temp(day, night, std=0, phase=False):
    return dict(day_temp=day, night_temp=night, temp_std=std, random_phase_offset=phase)


def _perlin(amp, period):
    return dict(amplitude_ratios=amp, nb_octaves=5, octaves_step=5, period=period)


_CAPS = {
    "no_noise": {10000: [10000], 15000: [15000]},
    "small_noise": {10000: [9000, 10000, 11000], 15000: [12500, 15000, 17500]},
    "big_noise": {10000: [7500, 9000, 10000, 11000, 12500], 15000: [10000, 12500, 15000, 17500, 20000]},
}

_CONFIG = {
    "default_house_prop": dict(id=1, init_air_temp=20, init_mass_temp=20, target_temp=20, deadband=0,
                               Ua=2.18e02, Cm=3.45e06, Ca=9.08e05, Hm=2.84e03, window_area=7.175,
                               shading_coeff=0.67, solar_gain_bool=True),
    "noise_house_prop": dict(noise_mode="big_start_temp", noise_parameters=_house_noise({
        "no_noise": (0, 0, 1, 1), "dwarf_noise": (0.05, 0.05, 1, 1), "house_small_noise": (0, 0, 0.9, 1.1),
        "house_medium_noise": (0, 0, 0.8, 1.2), "house_big_noise": (0, 0, 0.5, 1.5), "small_noise": (3, 1, 0.9, 1.1),
        "big_noise": (5, 2, 0.8, 1.2), "small_start_temp": (3, 0, 1, 1), "big_start_temp": (5, 0, 1, 1)})),
    "noise_house_prop_test": dict(noise_mode="small_start_temp", noise_parameters=_house_noise({
        "no_noise": (0, 0, 1, 1), "dwarf_noise": (0.05, 0.05, 1, 1), "small_noise": (3, 1, 0.9, 1.1),
        "big_noise": (5, 2, 0.8, 1.2), "small_start_temp": (3, 0, 1, 1), "big_start_temp": (5, 0, 1, 1)})),
    "default_hvac_prop": dict(id=1, COP=2.5, cooling_capacity=15000, latent_cooling_fraction=0.35,
                              lockout_duration=40, lockout_noise=0),
    "noise_hvac_prop": dict(noise_mode="no_noise", noise_parameters={
        k: dict(cooling_capacity_list=v) for k, v in _CAPS.items()}),
    # NB (reference quirk): the test HVAC noise has no cooling_capacity_list, so
    # MADemandResponseEnv(config, test=True) raises KeyError in the reference and here alike.
    "noise_hvac_prop_test": dict(noise_mode="no_noise", noise_parameters={
        "no_noise": dict(std_latent_cooling_fraction=0, factor_COP_low=1, factor_COP_high=1,
                         factor_cooling_capacity_low=1, factor_cooling_capacity_high=1),
        "small_noise": dict(std_latent_cooling_fraction=0.05, factor_COP_low=0.95, factor_COP_high=1.05,
                            factor_cooling_capacity_low=0.9, factor_cooling_capacity_high=1.1),
        "big_noise": dict(std_latent_cooling_fraction=0.1, factor_COP_low=0.85, factor_COP_high=1.15,
                          factor_cooling_capacity_low=0.6666667, factor_cooling_capacity_high=1.3333333333)}),
    "default_env_prop": {
        "start_datetime": "2021-01-01 00:00:00",
        "start_datetime_mode": "random",
        "time_step": 4,
        "cluster_prop": {
            "temp_mode": "noisy_sinusoidal_heatwave",
            "temp_parameters": {
                "constant": _temp(26.5, 26.5), "sinusoidal": _temp(30, 23), "sinusoidal_hot": _temp(30, 28),
                "sinusoidal_heatwave": _temp(34, 28), "sinusoidal_hot_heatwave": _temp(38, 32),
                "sinusoidal_cold_heatwave": _temp(30, 24), "sinusoidal_cold": _temp(24, 22),
                "noisy_sinusoidal": _temp(30, 23, 0.5), "noisy_sinusoidal_hot": _temp(30, 28, 0.5),
                "noisy_sinusoidal_heatwave": _temp(34, 28, 0.5), "noisier_sinusoidal_heatwave": _temp(34, 28, 2),
                "noisy_sinusoidal_cold": _temp(24, 22, 0.5), "shifting_sinusoidal": _temp(30, 23, 0, True),
                "shifting_sinusoidal_heatwave": _temp(34, 28, 0, True),
            },
            "nb_agents": 1,
            "nb_agents_comm": 10,
            "agents_comm_mode": "neighbours",
            "comm_defect_prob": 0,
            "agents_comm_parameters": {"neighbours_2D": {"row_size": 5, "distance_comm": 2}},
        },
        "state_properties": dict(hour=False, day=False, solar_gain=False, thermal=False, hvac=False),
        "message_properties": dict(thermal=False, hvac=False),
        "power_grid_prop": {
            "base_power_mode": "interpolation",
            "base_power_parameters": {
                "constant": dict(avg_power_per_hvac=4200, init_signal_per_hvac=910),
                "interpolation": dict(path_datafile="./monteCarlo/mergedGridSearchResultFinal.npy",
                                      path_parameter_dict="./monteCarlo/interp_parameters_dict.json",
                                      path_dict_keys="./monteCarlo/interp_dict_keys.csv",
                                      interp_update_period=300, interp_nb_agents=100),
            },
            "artificial_signal_ratio_range": 1,
            "artificial_ratio": 1.0,
            "signal_mode": "perlin",
            "signal_parameters": {
                "flat": {},
                "sinusoidals": dict(periods=[400, 1200], amplitude_ratios=[0.1, 0.3]),
                "regular_steps": dict(amplitude_per_hvac=6000, period=300),
                "perlin": _perlin(0.9, 400), "amplitude+_perlin": _perlin(0.9 * 1.1, 400),
                "amplitude++_perlin": _perlin(0.9 * 1.3, 400), "fast+_perlin": _perlin(0.9, 300),
                "fast++_perlin": _perlin(0.9, 200),
            },
        },
        "reward_prop": {
            "alpha_temp": 1, "alpha_sig": 1, "norm_reg_sig": 7500,
            "temp_penalty_mode": "individual_L2",
            "temp_penalty_parameters": {"individual_L2": {}, "common_L2": {}, "common_max_error": {},
                                        "mixture": dict(alpha_ind_L2=1, alpha_common_L2=1, alpha_common_max=0)},
            "sig_penalty_mode": "common_L2",
        },
    },
}

# grid of the Monte-Carlo base-power table (monteCarlo/interp_parameters_dict.json, key order of
# monteCarlo/interp_dict_keys.csv); the table itself is C-ordered over these axes
INTERP_KEYS = ("Ua_ratio", "Cm_ratio", "Ca_ratio", "Hm_ratio", "air_temp", "mass_temp", "OD_temp", "HVAC_power",
               "hour", "date")
INTERP_GRID = {
    "Ua_ratio": [0.9, 1, 1.1], "Cm_ratio": [0.9, 1, 1.1], "Ca_ratio": [0.9, 1, 1.1], "Hm_ratio": [0.9, 1, 1.1],
    "air_temp": [-4, -2, -1, -0.3, 0, 0.3, 1, 2, 4], "mass_temp": [-4, -2, 0, 2, 4],
    "OD_temp": [1, 3, 5, 7, 9, 11, 13, 15], "HVAC_power": [10000, 15000],
    "hour": [0.0, 10800.0, 21600.0, 25200.0, 27000.0, 39600.0, 46800.0, 57600.0, 61200.0, 63000.0, 75600.0, 86399.0],
    "date": [0, 79, 171, 263, 354, 364],
}


def default_config():
    """A fresh deep copy of the env-relevant part of the reference config."""
    return copy.deepcopy(_CONFIG)

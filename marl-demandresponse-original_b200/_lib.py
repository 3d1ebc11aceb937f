"""ctypes binding of include/mdr_b200.h.  There is no CPU fallback: if the CUDA library is
missing the import of the environment classes fails loudly."""
import ctypes as C
import os

from . import build as _build

MDR_ABI_VERSION = 10
MAX_SINUSOIDS, INTERP_DIMS, INTERP_MAX_AXIS, MAX_HOUSES_PER_ENV, MAX_HOUSES_PER_CLUSTER = 8, 10, 12, 1024, 16384
F32, F64 = 4, 8
COMM_NEIGHBOURS, COMM_TABLE, COMM_TABLE_PER_ENV, COMM_NONE = 0, 1, 2, 3
STATE_HOUR, STATE_DAY, STATE_SOLAR, STATE_THERMAL, STATE_HVAC = 1, 2, 4, 8, 16
MSG_THERMAL, MSG_HVAC = 1, 2
PEN = {"individual_L2": 0, "common_L2": 1, "common_max": 2, "mixture": 3}
BASE = {"constant": 0, "interpolation": 1}
SIG_FLAT, SIG_SINUSOIDALS, SIG_REGULAR_STEPS, SIG_PERLIN = 0, 1, 2, 3
ACT = {"array": 0, "bangbang": 1, "random": 2, "greedy": 3}
FLAG_NO_PIPELINE, FLAG_NO_FUSED, FLAG_NO_PDL, FLAG_NO_CLUSTER, FLAG_STATIC_TILES = 1, 2, 4, 8, 16
HAS_CLUSTER_PATH = True
METRIC_NAMES = ("steps", "sum_mean_reward", "sum_mean_temp_offset", "sum_mean_temp_error", "sum_sq_temp_error",
                "sum_sq_max_temp_error", "max_temp_error", "sum_od_temp", "sum_signal", "sum_consumption",
                "sum_signal_offset", "sum_signal_error", "sum_sq_signal_error")
N_METRICS = len(METRIC_NAMES)

_i32, _f64, _vp = C.c_int32, C.c_double, C.c_void_p


class MdrConfig(C.Structure):
    _fields_ = (
        [(n, _i32) for n in (
            "abi_version", "device", "precision", "n_envs", "n_houses", "n_comm", "n_features", "time_step",
            "comm_mode", "state_flags", "msg_flags", "temp_penalty_mode", "solar_gain", "base_power_mode",
            "signal_mode", "n_sinusoids", "interp_update_period", "interp_nb_agents", "perlin_nb_octaves",
            "perlin_octaves_step", "action_source", "obs_norm_agents")]
        + [(n, _f64) for n in (
            "alpha_temp", "alpha_sig", "norm_temp_penalty", "norm_sig_penalty", "mix_alpha_ind", "mix_alpha_common",
            "mix_alpha_max", "norm_reg_sig", "def_ua", "def_cm", "def_ca", "def_hm", "def_cop", "def_latent",
            "def_cap", "hvac_cop", "hvac_latent", "day_temp", "night_temp", "temp_std", "window_area",
            "shading_coeff", "avg_power_per_hvac")]
        + [("sin_periods", _f64 * MAX_SINUSOIDS), ("sin_ratios", _f64 * MAX_SINUSOIDS)]
        + [(n, _f64) for n in ("steps_amplitude_per_hvac", "steps_period", "perlin_amplitude", "perlin_period",
                               "comm_defect_prob")]
        + [("interp_dims", _i32 * INTERP_DIMS), ("interp_axes", (_f64 * INTERP_MAX_AXIS) * INTERP_DIMS),
           ("seed", C.c_uint64), ("l2_window_base", _vp), ("l2_window_bytes", C.c_uint64), ("l2_hit_ratio", _f64),
           ("flags", _i32), ("max_ctas", _i32)]
    )


class MdrHouses(C.Structure):
    _fields_ = [(n, _vp) for n in ("ua", "cm", "ca", "hm", "cap", "target", "deadband", "lockout_dur", "coef_a",
                                   "coef_b", "coef_c", "interp_key", "temps", "hvac")]


class MdrEnvs(C.Structure):
    _fields_ = [(n, _vp) for n in ("t_epoch", "phase", "od_temp", "solar_gain", "artificial_ratio",
                                   "max_power", "base_power", "signal", "cluster_power", "time_since_interp",
                                   "perlin_seed", "metrics", "workspace")]


class MdrStepInputs(C.Structure):
    _fields_ = [(n, _vp) for n in ("actions", "od_noise", "signal_noise", "interp_ids", "msg_keep", "comm_table",
                                   "interp_table")] + [("step_index", C.c_uint64), ("step_counter", _vp), ("env_mask", _vp)]


class MdrPopulationSpec(C.Structure):
    _fields_ = ([(n, _f64) for n in ("init_air_temp", "init_mass_temp", "target_temp", "deadband", "ua", "cm", "ca", "hm",
                                      "std_start_temp", "std_target_temp", "factor_thermo_low", "factor_thermo_high")]
                + [("cap_list", _f64 * 8), ("n_cap", _i32), ("lockout_duration", _i32), ("lockout_noise", _i32),
                   ("random_start", _i32), ("start_epoch", C.c_int64), ("random_phase", _i32),
                   ("interp_update_period", _i32), ("artificial_ratio", _f64), ("artificial_ratio_range", _f64)])


class MdrOutputs(C.Structure):
    _fields_ = [("obs", _vp), ("reward", _vp)]


EXPORTS = ("mdr_version", "mdr_strerror", "mdr_last_cuda_error", "mdr_obs_width", "mdr_validate",
           "mdr_launch_geometry", "mdr_precompute", "mdr_reset", "mdr_observe", "mdr_step", "mdr_step_host",
           "mdr_l2_persist_limit", "mdr_populate", "mdr_workspace_bytes", "mdr_sample_actions",
           "mdr_host_ctx_create", "mdr_host_ctx_destroy", "mdr_host_ctx_info")

_lib = None


class MdrError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """Loads (building first if the sources are newer) the CUDA library; raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MDR_LIB_PATH") or _build.LIB_PATH  # override: A/B-testing a differently built library
    if path == _build.LIB_PATH and build_if_missing and _build.is_stale():
        # build() serialises concurrent builders (torchrun ranks on a fresh clone) with a file lock and publishes
        # the library with an atomic rename, so no rank can dlopen a half-written file
        try:
            _build.build()
        except _build.NvccMissing as exc:  # no compiler on this box: the shipped .so is all there is
            if not os.path.isfile(path):
                raise MdrError("libmdr_b200.so is missing and could not be built: %s" % exc)
            import warnings
            warnings.warn("libmdr_b200.so is older than its sources and nvcc is not available to rebuild it; "
                          "loading the existing library (the ABI version is still checked)", RuntimeWarning)
        except Exception as exc:  # sources are newer and the build FAILED: never run a stale binary silently
            raise MdrError("libmdr_b200.so is stale and rebuilding it failed: %s" % exc)
    if not os.path.isfile(path):
        raise MdrError("CUDA extension %s not found; run `python __graft_entry__.py` (build) first. "
                       "There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise MdrError("libmdr_b200.so does not export %s" % name)
    lib.mdr_version.restype = C.c_int
    lib.mdr_strerror.restype = C.c_char_p
    lib.mdr_strerror.argtypes = [C.c_int]
    lib.mdr_last_cuda_error.restype = C.c_char_p
    P = C.POINTER
    lib.mdr_obs_width.argtypes = [P(MdrConfig)]
    lib.mdr_validate.argtypes = [P(MdrConfig)]
    lib.mdr_workspace_bytes.argtypes = [P(MdrConfig), P(C.c_size_t)]
    lib.mdr_sample_actions.argtypes = [_vp, C.c_int64, _i32, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp]
    lib.mdr_launch_geometry.argtypes = [P(MdrConfig), C.c_int, P(_i32), P(_i32), P(_i32), P(C.c_size_t), P(_i32), P(_i32)]
    lib.mdr_precompute.argtypes = [P(MdrConfig), P(MdrHouses), _vp]
    step_args = [P(MdrConfig), P(MdrHouses), P(MdrEnvs), P(MdrStepInputs), P(MdrOutputs)]
    lib.mdr_reset.argtypes = step_args + [_vp]
    lib.mdr_observe.argtypes = step_args + [_vp]
    lib.mdr_step.argtypes = step_args + [_i32, _vp]
    lib.mdr_step_host.argtypes = step_args + [_vp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.mdr_host_ctx_create.argtypes = [P(MdrConfig), _i32, _i32, P(_vp)]
    lib.mdr_host_ctx_destroy.argtypes = [_vp]
    lib.mdr_host_ctx_info.argtypes = [_vp, P(_i32), P(_i32), P(C.c_size_t)]
    lib.mdr_l2_persist_limit.argtypes = [C.c_int, C.c_size_t, P(C.c_size_t), P(C.c_size_t)]
    lib.mdr_populate.argtypes = [P(MdrConfig), P(MdrPopulationSpec), P(MdrHouses), P(MdrEnvs), _vp, C.c_uint64, _vp]
    if lib.mdr_version() != MDR_ABI_VERSION:
        raise MdrError("libmdr_b200.so ABI %d != binding ABI %d (rebuild)" % (lib.mdr_version(), MDR_ABI_VERSION))
    _lib = lib
    return lib


def check(status: int, what: str = "mdr call"):
    if status != 0:
        lib = load()
        msg = lib.mdr_strerror(status).decode()
        if status == -5:
            msg += ": " + lib.mdr_last_cuda_error().decode()
        if status == -3:
            raise ValueError("%s: %s" % (what, msg))
        raise MdrError("%s failed (%d): %s" % (what, status, msg))

"""ctypes binding of libsbx.so (C ABI declared in include/sbx.h).

There is no CPU fallback: if the shared library is missing or CUDA is not
usable, importing is fine but the first call raises `SbxLibraryError` loudly.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SBX_LIB: developer override used to A/B kernel variants (profiles/); same ABI, same checks
LIB_PATH = os.environ.get("SBX_LIB") or os.path.join(_HERE, "lib", "libsbx.so")

ABI_VERSION = 5
REWARD_REGRET, REWARD_ENERGY_CARBON = 0, 1
OPT_PIPELINE_CHUNKS = 1
OPT_L2_PREFETCH_DISTANCE = 2
OPT_HOST_SHARES = 3
OPT_NUMPY_MEANS = 4
OK = 0
MAX_ACTIONS = 3
MAX_HIST_BINS = 32
N_DEVICE_FIELDS = 15
DIAG_N = 16

ACT_BOILER_SETPOINT = 0
ACT_AHU_COOLING_SETPOINT = 1
ACT_AHU_HEATING_SETPOINT = 2
OBS_RAW = 0
OBS_HISTOGRAM = 1
PATH_AUTO = 0
PATH_STREAMING = 1
PATH_RESIDENT = 2
SOLVER_TF_JACOBI = 0
SOLVER_GAUSS_SEIDEL = 1
STEP_FIRST, STEP_MID, STEP_LAST = 0, 1, 2

# per-CV descriptor bits (include/sbx.h)
DESC_MATERIAL_SHIFT = 4
DESC_DIFFUSER = 0x0040
DESC_ZONE_SHIFT = 8
ZONE_NONE = 255

# field ids: (id, numpy dtype)
FIELDS = {
    "plan_desc": (1, np.uint16), "plan_material": (2, np.float64),
    "plan_cv_size": (3, np.float64), "zone_ncv": (4, np.int32),
    "zone_ndiff": (5, np.int32), "obs_zone_order": (6, np.int32),
    "reset_temps": (7, np.float32), "initial_temp": (8, np.float32),
    "ambient": (20, np.float64), "convection": (21, np.float64),
    "comfort": (22, np.uint8), "comfort_soon": (23, np.uint8),
    "occ_reward": (24, np.float64), "occ_obs": (25, np.int32),
    "occ_obs_zone": (30, np.float64),
    "price_elec": (26, np.float64), "carbon_elec": (27, np.float64),
    "price_gas": (28, np.float64), "time_features": (29, np.float64),
    "temp": (40, np.float32), "zone_mean": (41, np.float32),
    "global_mean": (42, np.float32), "q_cv": (43, np.float32),
    "thermostat_mode": (44, np.uint8), "ahu_heating_sp": (45, np.float64),
    "ahu_cooling_sp": (46, np.float64), "boiler_sp": (47, np.float64),
    "boiler_tank": (48, np.float64), "thermostat_prev": (49, np.int32),
    "episode": (50, np.int32), "convection_perm": (51, np.int32),
    "temp64": (52, np.float64), "q_cv64": (53, np.float64),
    "n_sweeps": (60, np.int32), "max_delta": (61, np.float32),
    "step_diag": (62, np.float64), "q_zone": (63, np.float64),
    "zone_supply_temp": (64, np.float64), "pre_zone_mean": (65, np.float32),
    "phase_cycles": (66, np.uint64),
}

DIAG = {
    "supply_air_temp": 0, "ahu_flow": 1, "boiler_flow": 2, "return_water": 3,
    "blower_w": 4, "ac_w": 5, "gas_w": 6, "pump_w": 7, "regret": 8,
    "norm_cost": 9, "norm_carbon": 10, "total_occ": 11, "productivity": 12,
    "cooling_requests": 13, "heating_requests": 14, "tank_temp": 15,
}


class SbxLibraryError(RuntimeError):
  pass


class SbxConfig(C.Structure):
  """Mirror of `sbx_config` (include/sbx.h); field order must match exactly."""
  _fields_ = [
      ("abi_version", C.c_int32),
      ("n_envs", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
      ("n_zones", C.c_int32), ("n_plans", C.c_int32), ("n_weather", C.c_int32),
      ("n_reset", C.c_int32), ("n_occ_zones", C.c_int32),
      ("n_table_steps", C.c_int32), ("episode_steps", C.c_int32),
      ("kernel_path", C.c_int32), ("solver", C.c_int32),
      ("iteration_limit", C.c_int32),
      ("time_step_sec", C.c_double), ("floor_height_m", C.c_double),
      ("convergence_threshold", C.c_double),
      ("comfort_heat", C.c_double), ("comfort_cool", C.c_double),
      ("eco_heat", C.c_double), ("eco_cool", C.c_double),
      ("ahu_recirculation", C.c_double), ("ahu_init_heating_setpoint", C.c_double),
      ("ahu_init_cooling_setpoint", C.c_double),
      ("ahu_fan_differential_pressure", C.c_double), ("ahu_fan_efficiency", C.c_double),
      ("ahu_max_air_flow_rate", C.c_double),
      ("boiler_init_setpoint", C.c_double), ("boiler_pump_head", C.c_double),
      ("boiler_pump_efficiency", C.c_double), ("boiler_heating_rate", C.c_double),
      ("boiler_cooling_rate", C.c_double), ("boiler_convection_coefficient", C.c_double),
      ("boiler_tank_length", C.c_double), ("boiler_tank_radius", C.c_double),
      ("boiler_water_capacity", C.c_double), ("boiler_insulation_conductivity", C.c_double),
      ("boiler_insulation_thickness", C.c_double),
      ("vav_max_air_flow_rate", C.c_double), ("vav_reheat_max_water_flow_rate", C.c_double),
      ("max_productivity_personhour_usd", C.c_double),
      ("min_productivity_personhour_usd", C.c_double),
      ("max_electricity_rate", C.c_double), ("max_natural_gas_rate", C.c_double),
      ("productivity_midpoint_delta", C.c_double),
      ("productivity_decay_stiffness", C.c_double),
      ("productivity_weight", C.c_double), ("energy_cost_weight", C.c_double),
      ("carbon_emission_weight", C.c_double), ("gas_carbon_rate", C.c_double),
      ("carbon_cost_factor", C.c_double), ("reward_normalizer_shift", C.c_double),
      ("reward_normalizer_scale", C.c_double), ("reward_kind", C.c_int32),
      ("reserved_reward", C.c_int32),
      ("discount_factor", C.c_double), ("occupancy_normalization_constant", C.c_double),
      ("n_actions", C.c_int32), ("action_target", C.c_int32 * MAX_ACTIONS),
      ("action_min", C.c_double * MAX_ACTIONS), ("action_max", C.c_double * MAX_ACTIONS),
      ("obs_mode", C.c_int32),
      ("obs_mean", C.c_double * N_DEVICE_FIELDS),
      ("obs_variance", C.c_double * N_DEVICE_FIELDS),
      ("n_hist_bins", C.c_int32 * 3),
      ("hist_bins", (C.c_double * MAX_HIST_BINS) * 3),
  ]


class SbxInfo(C.Structure):
  _fields_ = [
      ("obs_dim", C.c_int32), ("kernel_path", C.c_int32), ("n_sms", C.c_int32),
      ("resident_ctas_per_sm", C.c_int32), ("device_bytes", C.c_int64),
      ("kernel_launches", C.c_int64), ("sweeps_total", C.c_int64),
      ("env_steps_total", C.c_int64), ("step_count", C.c_int32),
      ("episode_ended", C.c_int32), ("time_index", C.c_int32), ("resident_kernel", C.c_int32),
  ]


class SbxTiming(C.Structure):
  _fields_ = [
      ("n_steps", C.c_int64), ("n_solve_launches", C.c_int64), ("step_ms", C.c_double),
      ("solve_ms", C.c_double), ("n_chunks", C.c_int32), ("reserved", C.c_int32),
  ]


EXPORTS = (
    "sbx_create", "sbx_destroy", "sbx_last_error", "sbx_get_info", "sbx_upload",
    "sbx_download", "sbx_reset", "sbx_step", "sbx_reset_host", "sbx_step_host",
    "sbx_fd_step", "sbx_sync", "sbx_abi_info", "sbx_host_alloc", "sbx_host_free",
    "sbx_set_option", "sbx_timing_begin", "sbx_timing_end", "sbx_set_device_convection",
)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
  """Loads libsbx.so; raises SbxLibraryError if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise SbxLibraryError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
        "g.build()'` (nvcc, sm_100a).  sbsim_b200 has no CPU fallback.")
  try:
    lib = C.CDLL(LIB_PATH)
  except OSError as e:
    raise SbxLibraryError(f"cannot load {LIB_PATH}: {e}") from e
  vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
  lib.sbx_create.argtypes = [C.POINTER(SbxConfig), C.c_int, C.POINTER(vp)]
  lib.sbx_destroy.argtypes = [vp]
  lib.sbx_last_error.argtypes = [vp]
  lib.sbx_last_error.restype = C.c_char_p
  lib.sbx_get_info.argtypes = [vp, C.POINTER(SbxInfo)]
  lib.sbx_upload.argtypes = [vp, C.c_int, vp, sz]
  lib.sbx_download.argtypes = [vp, C.c_int, vp, sz]
  lib.sbx_reset.argtypes = [vp, vp, vp, vp, vp, vp]
  lib.sbx_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
  lib.sbx_reset_host.argtypes = [vp, vp, vp, vp, vp]
  lib.sbx_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
  lib.sbx_fd_step.argtypes = [vp, vp, vp]
  lib.sbx_sync.argtypes = [vp]
  lib.sbx_abi_info.argtypes = [C.POINTER(C.c_int32), C.POINTER(sz), C.POINTER(sz)]
  lib.sbx_host_alloc.argtypes = [sz, C.POINTER(vp)]
  lib.sbx_host_free.argtypes = [vp]
  lib.sbx_set_option.argtypes = [vp, C.c_int, C.c_int64]
  lib.sbx_set_device_convection.argtypes = [vp, C.c_double, C.c_int32, C.c_uint64]
  lib.sbx_timing_begin.argtypes = [vp]
  lib.sbx_timing_end.argtypes = [vp, C.POINTER(SbxTiming)]
  for name in EXPORTS:
    if name != "sbx_last_error":
      getattr(lib, name).restype = C.c_int
  ver, cb, ib = C.c_int32(), sz(), sz()
  lib.sbx_abi_info(C.byref(ver), C.byref(cb), C.byref(ib))
  if (ver.value, cb.value, ib.value) != (ABI_VERSION, C.sizeof(SbxConfig), C.sizeof(SbxInfo)):
    raise SbxLibraryError(
        f"ABI mismatch: library (v{ver.value}, config {cb.value} B, info {ib.value} B) vs "
        f"binding (v{ABI_VERSION}, {C.sizeof(SbxConfig)} B, {C.sizeof(SbxInfo)} B); rebuild")
  _lib = lib
  return lib


class PinnedArray:
  """A NumPy array over page-locked host memory from sbx_host_alloc."""

  def __init__(self, shape, dtype):
    self._lib = load()
    self._ptr = C.c_void_p()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    rc = self._lib.sbx_host_alloc(nbytes, C.byref(self._ptr))
    if rc != OK:
      msg = self._lib.sbx_last_error(None)
      raise SbxLibraryError(f"sbx_host_alloc failed ({rc}): {msg.decode() if msg else ''}")
    buf = (C.c_char * max(nbytes, 1)).from_address(self._ptr.value)
    self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    self.array[...] = 0

  def __del__(self):
    try:
      if self._ptr.value:
        self._lib.sbx_host_free(self._ptr)
        self._ptr = C.c_void_p()
    except Exception:  # pylint: disable=broad-except
      pass


def pinned_time_step_arrays(batch: int, obs_dim: int):
  """(observation [B,D] f32, reward [B] f32, step_type [B] i32, discount [B] f32) as views
  of ONE page-locked block, in that order and back to back: sbx_step_host / sbx_reset_host
  then move all four outputs with a single device-to-host DMA."""
  block = PinnedArray((batch * (obs_dim + 3),), np.float32)
  flat = block.array
  o = batch * obs_dim
  views = (flat[:o].reshape(batch, obs_dim), flat[o:o + batch],
           flat[o + batch:o + 2 * batch].view(np.int32), flat[o + 2 * batch:o + 3 * batch])
  return block, views


class Handle:
  """Owns one `sbx_handle`.  Raises SbxLibraryError with the library's message."""

  def __init__(self, cfg: SbxConfig, device: int = 0):
    self._lib = load()
    self.cfg = cfg
    self._h = C.c_void_p()
    rc = self._lib.sbx_create(C.byref(cfg), int(device), C.byref(self._h))
    if rc != OK:
      msg = self._lib.sbx_last_error(None)
      raise SbxLibraryError(f"sbx_create failed ({rc}): {msg.decode() if msg else ''}")

  def _check(self, rc: int, what: str):
    if rc != OK:
      msg = self._lib.sbx_last_error(self._h)
      raise SbxLibraryError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

  def close(self):
    if getattr(self, "_h", None) is not None and self._h.value:
      self._lib.sbx_destroy(self._h)
      self._h = C.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass

  def info(self) -> SbxInfo:
    out = SbxInfo()
    self._check(self._lib.sbx_get_info(self._h, C.byref(out)), "sbx_get_info")
    return out

  def upload(self, name: str, array) -> None:
    fid, dtype = FIELDS[name]
    a = np.ascontiguousarray(array, dtype=dtype)
    self._check(self._lib.sbx_upload(self._h, fid, a.ctypes.data, a.nbytes),
                f"sbx_upload({name})")

  def download(self, name: str, shape) -> np.ndarray:
    fid, dtype = FIELDS[name]
    out = np.empty(shape, dtype=dtype)
    self._check(self._lib.sbx_download(self._h, fid, out.ctypes.data, out.nbytes),
                f"sbx_download({name})")
    return out

  def reset_device(self, obs_ptr, reward_ptr, step_type_ptr, discount_ptr, stream=0):
    self._check(self._lib.sbx_reset(self._h, obs_ptr, reward_ptr, step_type_ptr,
                                    discount_ptr, stream), "sbx_reset")

  def step_device(self, action_ptr, obs_ptr, reward_ptr, step_type_ptr, discount_ptr,
                  stream=0):
    self._check(self._lib.sbx_step(self._h, action_ptr, obs_ptr, reward_ptr,
                                   step_type_ptr, discount_ptr, stream), "sbx_step")

  def reset_host(self, obs, reward, step_type, discount):
    self._check(self._lib.sbx_reset_host(self._h, obs.ctypes.data, reward.ctypes.data,
                                         step_type.ctypes.data, discount.ctypes.data),
                "sbx_reset_host")

  def step_host(self, action, obs, reward, step_type, discount):
    self._check(self._lib.sbx_step_host(self._h, action.ctypes.data, obs.ctypes.data,
                                        reward.ctypes.data, step_type.ctypes.data,
                                        discount.ctypes.data), "sbx_step_host")

  def fd_step(self, ambient, convection):
    a = np.ascontiguousarray(ambient, dtype=np.float64)
    c = np.ascontiguousarray(convection, dtype=np.float64)
    self._check(self._lib.sbx_fd_step(self._h, a.ctypes.data, c.ctypes.data),
                "sbx_fd_step")

  def sync(self):
    self._check(self._lib.sbx_sync(self._h), "sbx_sync")

  def set_option(self, option: int, value: int):
    self._check(self._lib.sbx_set_option(self._h, int(option), C.c_int64(int(value))),
                "sbx_set_option")

  def set_device_convection(self, p: float, distance: int, seed: int):
    self._check(self._lib.sbx_set_device_convection(
        self._h, C.c_double(float(p)), C.c_int32(int(distance)), C.c_uint64(int(seed) & (2**64 - 1))),
        "sbx_set_device_convection")

  def timing_begin(self):
    self._check(self._lib.sbx_timing_begin(self._h), "sbx_timing_begin")

  def timing_end(self) -> SbxTiming:
    out = SbxTiming()
    self._check(self._lib.sbx_timing_end(self._h, C.byref(out)), "sbx_timing_end")
    return out

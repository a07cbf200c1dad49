"""Batched, GPU-resident replacement for sbsim's Environment + SimulatorBuilding.

`Environment` keeps the reference's TF-Agents PyEnvironment surface
(`reset / step / current_time_step / observation_spec / action_spec /
time_step_spec / batched / batch_size`; environment/environment.py:352,
1165-1368 of the reference) for B independent buildings at once: every array of
the returned TimeStep gains a leading batch dimension, exactly what TF-Agents
expects from a batched py_environment.

`SimulatorBuilding` stands where the reference's SimulatorBuilding +
TFSimulator + FloorPlanBasedBuilding stand (simulator/simulator_building.py:44,
simulator/tf_simulator.py:502, simulator/building.py:609): it owns the compiled
floor plans, the HVAC parameters, the weather and the simulation clock.  All
per-step arithmetic happens in libsbx.so (hand-written sm_100a CUDA) through the
C ABI of include/sbx.h; this module only builds configuration and tables and
moves I/O buffers.  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import dataclasses
import os
import warnings
from typing import List, Optional, Sequence, Union

import numpy as np
import pandas as pd

from sbsim_b200 import _lib
from sbsim_b200 import config as config_lib
from sbsim_b200 import exogenous
from sbsim_b200 import floorplan
from sbsim_b200 import specs

ACTION_REJECTION_REWARD = -np.inf  # environment.py:52 (never produced: simulated devices accept every in-range action)

AHU_FIELDS = (
    "cooling_request_count", "differential_pressure_setpoint",
    "discharge_fan_speed_percentage_command", "outside_air_flowrate_sensor",
    "outside_air_temperature_sensor", "supply_air_cooling_temperature_setpoint",
    "supply_air_flowrate_sensor", "supply_air_heating_temperature_setpoint",
    "supply_fan_speed_percentage_command")
BOILER_FIELDS = ("heating_request_count", "supply_water_setpoint",
                 "supply_water_temperature_sensor")
VAV_FIELDS = ("supply_air_damper_percentage_command", "supply_air_flowrate_setpoint",
              "zone_air_temperature_sensor")
# action-settable setpoints in `building.devices` order (simulator_building.py:71-74)
_ACTION_SETPOINTS = (
    ("boiler", "supply_water_setpoint", _lib.ACT_BOILER_SETPOINT),
    ("air_handler", "supply_air_cooling_temperature_setpoint", _lib.ACT_AHU_COOLING_SETPOINT),
    ("air_handler", "supply_air_heating_temperature_setpoint", _lib.ACT_AHU_HEATING_SETPOINT),
)


class SimulatorBuilding:
  """B simulated buildings: plans + HVAC + weather + occupancy + clock."""

  def __init__(self,
               plans: Union[floorplan.CompiledPlan, Sequence[floorplan.CompiledPlan]],
               hvac: config_lib.FloorPlanBasedHvac,
               weather_controller,
               occupancy,
               *,
               n_envs: Optional[int] = None,
               time_step_sec: float = 300.0,
               convergence_threshold: float = 0.1,
               iteration_limit: int = 100,
               iteration_warning: int = 30,
               start_timestamp: pd.Timestamp,
               floor_height_cm: float = 300.0,
               initial_temp: Union[float, np.ndarray] = 294.0,
               reset_temp_values: Optional[np.ndarray] = None,
               convection_simulator=None,
               solver: str = "tf_jacobi"):
    # The reference picks the diffusion model by simulator class: TFSimulator
    # (fp32 Jacobi, tf_simulator.py:502; calibrated config) or
    # SimulatorFlexibleGeometries (fp64 Gauss-Seidel, simulator.py:278; legacy config).
    if solver not in ("tf_jacobi", "gauss_seidel"):
      raise ValueError("solver must be 'tf_jacobi' or 'gauss_seidel'")
    self.solver = solver
    self.convection_simulator = convection_simulator
    if isinstance(plans, floorplan.CompiledPlan):
      plans = [plans]
    self.plans: List[floorplan.CompiledPlan] = list(plans)
    if n_envs is None:
      n_envs = len(self.plans)
    if len(self.plans) not in (1, n_envs):
      raise ValueError("pass one plan (shared by every env) or one plan per env")
    self.n_envs = int(n_envs)
    self.hvac = hvac
    # The deprecated Hvac (hvac.py:35-123) addresses zones by (row, column) and names its devices
    # `vav_<i>_<j>`: the rooms of a legacy_building() plan (raster order = row-major) take those
    # names, and the observation order follows the new device ids.
    self._legacy_coordinates = None
    if isinstance(hvac, config_lib.Hvac):
      coords = sorted(hvac.zone_coordinates)
      renamed = []
      for pl in self.plans:
        if len(coords) != pl.n_zones:
          raise ValueError(f"Hvac has {len(coords)} zone coordinates, the building {pl.n_zones} rooms; build "
                           "the plan with legacy_building(room_shape, building_shape)")
        names = [f"{i}_{j}" for i, j in coords]
        order = np.array(sorted(range(len(names)), key=lambda k: "vav_" + names[k]), dtype=np.int32)
        renamed.append(dataclasses.replace(pl, zone_names=names, obs_zone_order=order))
      self.plans = renamed
      self._legacy_coordinates = {f"{i}_{j}": (i, j) for i, j in coords}
    if isinstance(weather_controller, (list, tuple)):
      if len(weather_controller) != self.n_envs:
        raise ValueError("pass one weather controller or one per env")
    self.weather_controller = weather_controller
    self.occupancy = occupancy
    self.time_step_sec = float(time_step_sec)
    self.convergence_threshold = float(convergence_threshold)
    self.iteration_limit = int(iteration_limit)
    self.iteration_warning = int(iteration_warning)
    self.start_timestamp = start_timestamp
    self.floor_height_cm = float(floor_height_cm)
    self.initial_temp = initial_temp
    self.reset_temp_values = reset_temp_values
    self.height, self.width = self.plans[0].height, self.plans[0].width
    self.n_zones = max(1, max(p.n_zones for p in self.plans))

  @property
  def zone_ids(self) -> List[str]:
    """zone ids of plan 0 (conversion_utils.floor_plan_based_zone_identifier_to_id; with the
    deprecated Hvac: conversion_utils.zone_coordinates_to_id, 'zone_id_(i, j)')."""
    if self._legacy_coordinates is not None:
      return ["zone_id_" + str(self._legacy_coordinates[n]) for n in self.plans[0].zone_names]
    return ["zone_id_" + n.replace("room_", "") for n in self.plans[0].zone_names]


class BatchedWeather:
  """Per-building ambient series given directly as a table [B, T] (used for the
  randomised-weather configs); `table(timestamps)` must return that shape."""

  def __init__(self, table_fn, convection_coefficient: np.ndarray):
    self._table_fn = table_fn
    self.convection_coefficient = np.asarray(convection_coefficient, dtype=np.float64)

  def table(self, timestamps):
    return self._table_fn(timestamps)


class Environment:
  """Batched building-control environment (TF-Agents PyEnvironment surface)."""

  def __init__(self,
               building: SimulatorBuilding,
               reward_function: config_lib.SetpointEnergyCarbonRegretFunction,
               observation_normalizer: config_lib.StandardScoreObservationNormalizer,
               action_config: config_lib.ActionConfig,
               discount_factor: float = 1,
               metrics_path: Optional[str] = None,
               num_days_in_episode: float = 3,
               default_actions=None,
               label: str = "episode_metrics",
               num_hod_features: int = 1,
               num_dow_features: int = 1,
               occupancy_normalization_constant: float = 0.0,
               observation_histogram_reducer: Optional[config_lib.HistogramReducer] = None,
               time_zone: str = "US/Pacific",
               step_interval: pd.Timedelta = pd.Timedelta(5, unit="minutes"),
               writer_factory=None,
               device_action_tuples=None,
               metrics_reporting_interval: float = 100,
               run_command_predictors=None,
               image_generator=None,
               *,
               device: int = 0,
               kernel_path: int = _lib.PATH_AUTO,
               metrics_env_indices: Sequence[int] = (0,),
               numpy_zone_means: bool = False):
    if discount_factor <= 0 or discount_factor > 1:
      raise ValueError("Discount factor must be in (0,1]")   # environment.py:445-446
    if num_hod_features != 1 or num_dow_features != 1:
      raise NotImplementedError("one hod and one dow feature pair, as in sim_config.gin:597-598")
    # Constructor arguments of the reference (environment.py:355-383) that have no batched
    # counterpart are accepted, so that a call site written for the reference keeps working, and
    # reported once: the action set comes from `action_config` (all envs share it), TensorBoard
    # summaries / run-command predictors / building images are host-side tooling around the path.
    for name, value in (("device_action_tuples", device_action_tuples),
                        ("run_command_predictors", run_command_predictors),
                        ("image_generator", image_generator)):
      if value:
        warnings.warn(f"sbsim_b200.Environment ignores {name}", stacklevel=2)
    del metrics_reporting_interval      # TensorBoard reporting period of the reference: nothing to report to
    self.building = building
    self.reward_function = reward_function
    self.discount_factor = float(discount_factor)
    self._label = label
    self._time_zone = time_zone
    # Episode logs (environment.py:1181-1198): with `metrics_path`, every reset() opens
    # <metrics_path>/<label>_<yymmdd_HHMMSS>/ and step() logs the buildings `metrics_env_indices`
    # in the reference ProtoWriter's shard format (sbsim_b200/episode_writer.py).  `writer_factory`
    # of the reference produces writers that take protobuf messages; here the records are
    # serialized directly, so a factory is accepted but only its presence is honoured.
    self._metrics_path = metrics_path
    self._writer_factory = writer_factory
    self._metrics_env_indices = tuple(int(i) for i in metrics_env_indices)
    self._metrics_writer = None
    self._occupancy_normalization_constant = float(occupancy_normalization_constant)
    self._observation_normalizer = observation_normalizer
    self._observation_histogram_reducer = observation_histogram_reducer
    self._start_timestamp = building.start_timestamp
    self._end_timestamp = self._start_timestamp + pd.Timedelta(num_days_in_episode, unit="days")
    self._step_interval = step_interval
    self._num_timesteps_in_episode = int(
        (self._end_timestamp - self._start_timestamp) / self._step_interval)
    self._episode_ended = False
    self._step_count = 0
    self._time_index = 0
    self._current_time_step = None
    self._device = int(device)

    # ---- action spec (environment.py:591-653) ----
    self._action_names: List[str] = []
    self._action_fields: List[tuple] = []      # (device, bare setpoint name) per action dimension
    self._action_normalizers = {}
    targets, lo, hi = [], [], []
    for dev, setpoint, target in _ACTION_SETPOINTS:
      norm = action_config.get_action_normalizer(setpoint)
      if not norm:
        continue
      field_id = f"{dev}_{setpoint}"
      self._action_names.append(field_id)
      self._action_fields.append((dev, setpoint))
      self._action_normalizers[field_id] = norm
      targets.append(target)
      lo.append(norm.setpoint_min)
      hi.append(norm.setpoint_max)
    self._action_spec = specs.BoundedArraySpec(
        shape=(len(self._action_names),), dtype=np.float32, name="action",
        minimum=-1.0, maximum=1.0)

    # ---- observation spec (environment.py:698-812) ----
    b = building
    names = [f"air_handler_{f}" for f in AHU_FIELDS] + [f"boiler_{f}" for f in BOILER_FIELDS]
    hist = observation_histogram_reducer
    if hist is None:
      plan0 = b.plans[0]
      for slot in range(b.n_zones):
        zname = (plan0.zone_names[plan0.obs_zone_order[slot]]
                 if slot < plan0.n_zones else f"unused_{slot}")
        names += [f"vav_{zname}_{f}" for f in VAV_FIELDS]
    else:
      for f in VAV_FIELDS:
        if f not in hist.histogram_parameters:
          raise NotImplementedError(
              f"histogram mode needs bins for every VAV measurement; missing {f}")
        names += [f"{f}_h_{v:.2f}" for v in hist.histogram_parameters[f]]
    names += ["hod_cos_000", "hod_sin_000", "dow_cos_000", "dow_sin_000",
              "comfort_mode_now", "comfort_mode_soon", "num_occupants"]
    self._field_names = names
    self._observation_spec = specs.ArraySpec(shape=(len(names),), dtype=np.float32,
                                             name="observation")

    # ---- C config ----
    cfg = _lib.SbxConfig()
    cfg.abi_version = _lib.ABI_VERSION
    cfg.n_envs, cfg.height, cfg.width, cfg.n_zones = b.n_envs, b.height, b.width, b.n_zones
    cfg.n_plans = len(b.plans)
    per_env_weather = isinstance(b.weather_controller, (list, tuple, BatchedWeather))
    cfg.n_weather = b.n_envs if per_env_weather else 1
    if b.reset_temp_values is not None:
      rt = np.asarray(b.reset_temp_values)
      cfg.n_reset = 1 if rt.ndim == 2 else rt.shape[0]
    else:
      cfg.n_reset = 0
    per_zone_occ = bool(getattr(b.occupancy, "per_zone", False)) or (
        isinstance(b.occupancy, exogenous.TableOccupancy)
        and b.occupancy.reward_table.shape[1] > 1)
    cfg.n_occ_zones = b.n_zones if per_zone_occ else 1
    cfg.episode_steps = self._num_timesteps_in_episode
    # rows 0 .. N+1: an episode is N+1 steps (the terminal one included), and the
    # reference never evaluates weather / occupancy beyond t_{N+1} either
    cfg.n_table_steps = self._num_timesteps_in_episode + 2
    cfg.kernel_path = int(kernel_path)
    cfg.solver = (_lib.SOLVER_GAUSS_SEIDEL if b.solver == "gauss_seidel"
                  else _lib.SOLVER_TF_JACOBI)
    cfg.time_step_sec = b.time_step_sec
    cfg.floor_height_m = b.floor_height_cm / 100.0
    cfg.convergence_threshold = b.convergence_threshold
    cfg.iteration_limit = b.iteration_limit
    sched = b.hvac.schedule
    cfg.comfort_heat, cfg.comfort_cool = map(float, sched.comfort_temp_window)
    cfg.eco_heat, cfg.eco_cool = map(float, sched.eco_temp_window)
    ah, bl = b.hvac.air_handler, b.hvac.boiler
    cfg.ahu_recirculation = ah.recirculation
    cfg.ahu_init_heating_setpoint = ah.heating_air_temp_setpoint
    cfg.ahu_init_cooling_setpoint = ah.cooling_air_temp_setpoint
    cfg.ahu_fan_differential_pressure = ah.fan_differential_pressure
    cfg.ahu_fan_efficiency = ah.fan_efficiency
    cfg.ahu_max_air_flow_rate = ah.max_air_flow_rate
    cfg.boiler_init_setpoint = bl.reheat_water_setpoint
    cfg.boiler_pump_head = bl.water_pump_differential_head
    cfg.boiler_pump_efficiency = bl.water_pump_efficiency
    cfg.boiler_heating_rate = bl.heating_rate
    cfg.boiler_cooling_rate = bl.cooling_rate
    cfg.boiler_convection_coefficient = bl.convection_coefficient
    cfg.boiler_tank_length = bl.tank_length
    cfg.boiler_tank_radius = bl.tank_radius
    cfg.boiler_water_capacity = bl.water_capacity
    cfg.boiler_insulation_conductivity = bl.insulation_conductivity
    cfg.boiler_insulation_thickness = bl.insulation_thickness
    cfg.vav_max_air_flow_rate = b.hvac.vav_max_air_flow_rate
    cfg.vav_reheat_max_water_flow_rate = b.hvac.vav_reheat_max_water_flow_rate
    rf = reward_function
    if isinstance(rf, config_lib.SetpointEnergyCarbonRewardFunction):
      cfg.reward_kind = _lib.REWARD_ENERGY_CARBON
      cfg.max_productivity_personhour_usd = rf.max_productivity_personhour_usd
      cfg.productivity_midpoint_delta = rf.productivity_midpoint_delta
      cfg.productivity_decay_stiffness = rf.productivity_decay_stiffness
      cfg.energy_cost_weight = rf.energy_cost_weight
      cfg.carbon_emission_weight = rf.carbon_cost_weight
      cfg.carbon_cost_factor = rf.carbon_cost_factor
      cfg.reward_normalizer_shift = rf.reward_normalizer_shift
      cfg.reward_normalizer_scale = rf.reward_normalizer_scale
    else:
      cfg.reward_kind = _lib.REWARD_REGRET
      cfg.max_productivity_personhour_usd = rf.max_productivity_personhour_usd
      cfg.min_productivity_personhour_usd = rf.min_productivity_personhour_usd
      cfg.max_electricity_rate = rf.max_electricity_rate
      cfg.max_natural_gas_rate = rf.max_natural_gas_rate
      cfg.productivity_midpoint_delta = rf.productivity_midpoint_delta
      cfg.productivity_decay_stiffness = rf.productivity_decay_stiffness
      cfg.productivity_weight = rf.productivity_weight
      cfg.energy_cost_weight = rf.energy_cost_weight
      cfg.carbon_emission_weight = rf.carbon_emission_weight
    cfg.gas_carbon_rate = rf.natural_gas_energy_cost.carbon_rate
    cfg.discount_factor = self.discount_factor
    cfg.occupancy_normalization_constant = self._occupancy_normalization_constant
    cfg.n_actions = len(targets)
    for i, (t, a, z) in enumerate(zip(targets, lo, hi)):
      cfg.action_target[i], cfg.action_min[i], cfg.action_max[i] = t, a, z
    cfg.obs_mode = _lib.OBS_RAW if hist is None else _lib.OBS_HISTOGRAM
    for i, f in enumerate(AHU_FIELDS + BOILER_FIELDS + VAV_FIELDS):
      cfg.obs_mean[i], cfg.obs_variance[i] = observation_normalizer.get(f)
    if hist is not None:
      for i, f in enumerate(VAV_FIELDS):
        bins = hist.histogram_parameters[f]
        if len(bins) > _lib.MAX_HIST_BINS:
          raise ValueError(f"at most {_lib.MAX_HIST_BINS} histogram bins per measurement")
        cfg.n_hist_bins[i] = len(bins)
        for j, v in enumerate(bins):
          cfg.hist_bins[i][j] = float(v)
    self._cfg = cfg
    self._handle = _lib.Handle(cfg, self._device)
    info = self._handle.info()
    if info.obs_dim != len(names):
      raise RuntimeError(f"observation size mismatch: library {info.obs_dim}, host {len(names)}")
    self.kernel_path = info.kernel_path
    # Zone / grid means summed in the order of the reference's np.mean (SBX_OPT_NUMPY_MEANS):
    # bit-identical to building.py:845-871 instead of the exact integer mean a few ulp away.
    self.numpy_zone_means = bool(numpy_zone_means)
    if self.numpy_zone_means:
      self._handle.set_option(_lib.OPT_NUMPY_MEANS, 1)
    self._upload_static()
    self._upload_tables()
    B, D = b.n_envs, len(names)
    # Outputs land in page-locked host buffers (direct DMA, no staging memcpy).
    # Two sets alternate, so the arrays of a returned TimeStep stay valid until the
    # next-but-one reset()/step() call; copy them to keep them longer.
    # two alternating sets; each set is ONE pinned block so a step's outputs come back
    # in a single DMA (sbx_step_host recognises back-to-back buffers)
    self._pinned_blocks, self._pinned = [], []
    for _ in range(2):
      block, views = _lib.pinned_time_step_arrays(B, D)
      self._pinned_blocks.append(block)
      self._pinned.append(views)
    self._pinned_action = _lib.PinnedArray((B, max(len(targets), 1)), np.float32)
    self._flip = 0
    self._bind_outputs()
    # stochastic convection: exact host replay, one generator stream per env
    self._conv = b.convection_simulator
    if self._conv is not None and getattr(self._conv, "mode", "replay") == "device":
      # device-RNG mode: nothing is drawn on the host (see sbsim_b200/convection.py)
      seed = self._conv._seed if self._conv._seed is not None else int.from_bytes(__import__("os").urandom(8), "little")
      self._handle.set_device_convection(self._conv._p, self._conv._distance, seed)
      self._conv = None
    if self._conv is not None:
      self._conv_streams = [self._conv.make_stream() for _ in range(B)]
      self._conv_rooms = []
      for plan in b.plans:
        rooms = []
        for zi in range(plan.n_zones):
          r, c = plan.zone_indices(zi)
          rooms.append(list(zip(r.tolist(), c.tolist())))
        self._conv_rooms.append(rooms)

  # ---- uploads -------------------------------------------------------------

  def _upload_static(self):
    b, h = self.building, self._handle
    packed = floorplan.pack_plans(b.plans, n_zones=b.n_zones)
    for k, v in packed.items():
      h.upload(k, v)
    if b.reset_temp_values is not None:
      rt = np.asarray(b.reset_temp_values, dtype=np.float32)
      if rt.shape[-2:] != (b.height, b.width):
        raise ValueError(f"reset_temp_values shape {rt.shape} != grid {(b.height, b.width)}")
      h.upload("reset_temps", rt)
    it = np.broadcast_to(np.asarray(b.initial_temp, dtype=np.float32), (b.n_envs,))
    h.upload("initial_temp", it)

  def _upload_tables(self):
    b, h = self.building, self._handle
    T = self._cfg.n_table_steps
    ts = exogenous.step_timestamps(self._start_timestamp, b.time_step_sec, T)
    self._timestamps = ts
    w = b.weather_controller
    if isinstance(w, BatchedWeather):
      amb = np.asarray(w.table(ts), dtype=np.float64)
      conv = np.broadcast_to(w.convection_coefficient, (b.n_envs,))
    elif isinstance(w, (list, tuple)):
      amb = np.stack([wi.table(ts) for wi in w])
      conv = np.array([wi.get_air_convection_coefficient(ts[0]) for wi in w])
    else:
      amb = w.table(ts)[None, :]
      conv = np.array([w.get_air_convection_coefficient(ts[0])])
    if amb.shape != (self._cfg.n_weather, T):
      raise ValueError(f"ambient table shape {amb.shape} != {(self._cfg.n_weather, T)}")
    h.upload("ambient", amb)
    h.upload("convection", conv)
    self._ambient_table = amb            # host copies, read by integration/cuda_simulator.py
    sched = b.hvac.schedule
    h.upload("comfort", sched.table(ts))
    soon = pd.Timedelta(60, unit="minute")
    h.upload("comfort_soon", sched.table([t + soon for t in ts]))
    self._upload_occupancy()
    pe, ce, pg = exogenous.energy_tables(self.reward_function.electricity_energy_cost,
                                         self.reward_function.natural_gas_energy_cost, ts)
    h.upload("price_elec", pe)
    h.upload("carbon_elec", ce)
    h.upload("price_gas", pg)
    h.upload("time_features", exogenous.time_feature_table(ts))

  def _upload_occupancy(self):
    """Occupancy tables of ONE episode.  A stateful model (the shipped
    RandomizedArrivalDepartureOccupancy) keeps drawing from its generator, so the
    tables are rebuilt at every reset after the first, continuing the stream like
    the reference does over consecutive (complete) episodes."""
    b, h = self.building, self._handle
    T = self._cfg.n_table_steps
    occ_r, occ_o, occ_z = exogenous.occupancy_tables(
        b.occupancy, b.zone_ids or ["zone_id_0"], self._timestamps, b.time_step_sec,
        per_zone=self._cfg.n_occ_zones > 1)
    if occ_r.shape[0] < T or occ_o.shape[0] < T:
      raise ValueError("occupancy tables shorter than the episode")
    h.upload("occ_reward", occ_r[:T])
    h.upload("occ_obs", occ_o[:T])
    self._occ_reward_table, self._occ_obs_table = occ_r[:T], occ_o[:T]
    if occ_z is not None:      # per-building count of the zones each plan really has
      h.upload("occ_obs_zone", occ_z[:T])
    self._occupancy_episodes = getattr(self, "_occupancy_episodes", 0) + 1

  def _before_reset(self):
    if getattr(self.building.occupancy, "stateful", False) and self._current_time_step is not None:
      self._upload_occupancy()

  # ---- PyEnvironment surface ----------------------------------------------

  @property
  def batched(self) -> bool:
    return True

  @property
  def batch_size(self) -> int:
    return self.building.n_envs

  def action_spec(self):
    return self._action_spec

  def observation_spec(self):
    return self._observation_spec

  def time_step_spec(self):
    return specs.TimeStep(
        step_type=specs.ArraySpec((), np.int32, "step_type"),
        reward=specs.ArraySpec((), np.float32, "reward"),
        discount=specs.BoundedArraySpec((), np.float32, 0.0, 1.0, "discount"),
        observation=self._observation_spec)

  def current_time_step(self):
    return self._current_time_step

  @property
  def field_names(self) -> List[str]:
    return list(self._field_names)

  @property
  def action_names(self) -> List[str]:
    return list(self._action_names)

  @property
  def label(self) -> str:
    return self._label

  @property
  def steps_per_episode(self) -> int:
    return int((self._end_timestamp - self._start_timestamp).total_seconds()
               // self.building.time_step_sec)

  @property
  def start_timestamp(self) -> pd.Timestamp:
    return self._start_timestamp

  @property
  def end_timestamp(self) -> pd.Timestamp:
    return self._end_timestamp

  @property
  def current_simulation_timestamp(self) -> pd.Timestamp:
    return self._start_timestamp + self._time_index * pd.Timedelta(
        self.building.time_step_sec, unit="s")

  @property
  def handle(self) -> _lib.Handle:
    return self._handle

  def _bind_outputs(self):
    self._obs, self._reward, self._step_type, self._discount = self._pinned[self._flip]

  def _time_step(self) -> specs.TimeStep:
    ts = specs.TimeStep(step_type=self._step_type, reward=self._reward,
                        discount=self._discount, observation=self._obs)
    self._flip ^= 1
    self._bind_outputs()
    return ts

  def _sync_counters(self):
    info = self._handle.info()
    self._step_count = info.step_count
    self._episode_ended = bool(info.episode_ended)
    self._time_index = info.time_index

  def reset(self) -> specs.TimeStep:
    """Environment._reset (environment.py:1165-1212) for every env."""
    self._before_reset()
    self._handle.reset_host(self._obs, self._reward, self._step_type, self._discount)
    self._episode_ended = False
    self._step_count = 0
    self._time_index = 0
    self._current_time_step = self._time_step()
    if self._metrics_path:
      from sbsim_b200 import episode_writer      # pylint: disable=g-import-not-at-top
      now = pd.Timestamp.now("UTC")
      out = os.path.join(self._metrics_path, f"{self._label}_{now:%y%m%d_%H%M%S}")
      self._metrics_writer = episode_writer.EpisodeWriter(self, out, self._metrics_env_indices)
    return self._current_time_step

  def _validate_action(self, action) -> np.ndarray:
    a = np.ascontiguousarray(action, dtype=np.float32)
    want = (self.batch_size, len(self._action_names))
    if a.shape != want:
      raise ValueError(f"action shape {a.shape} != {want}")
    return a

  @staticmethod
  def _raise_action_bounds(a: np.ndarray):
    """bounded_action_normalizer.py:84-90; the range check itself runs inside sbx_step_host (one
    pass in C before anything is stepped), this formats the reference's error."""
    tol = config_lib.ACTION_TOLERANCE
    bad = a[(a < -1.0 - tol) | (a > 1.0 + tol) | np.isnan(a)][0]
    raise ValueError(f"agent_action: {bad} not within bounds [-1.0, 1.0]")

  def step(self, action) -> specs.TimeStep:
    """Environment._step (environment.py:1228-1360) for every env.

    The returned arrays are views of page-locked buffers that the library fills by
    DMA; two buffer sets alternate, so a TimeStep stays valid until the
    next-but-one reset()/step().  Copy it to keep it longer."""
    if self._current_time_step is None:       # PyEnvironment.step [TF-Agents]
      return self.reset()
    if self._episode_ended:                   # environment.py:1252-1253
      return self.reset()
    a = self._validate_action(action)
    if self._conv is not None:
      # the host-drawn convection permutation advances the reference's generators: actions that
      # will be refused must not get that far
      tol = config_lib.ACTION_TOLERANCE
      if a.size and not (np.abs(a).max() <= 1.0 + tol):
        self._raise_action_bounds(a)
      self._upload_convection()
    try:      # pageable actions are staged by the library; it also checks their bounds
      self._handle.step_host(a if a.size else self._pinned_action.array,
                             self._obs, self._reward, self._step_type, self._discount)
    except _lib.SbxLibraryError as e:
      if "not within bounds" in str(e):
        self._raise_action_bounds(a)
      raise
    ended = self._step_count >= self._num_timesteps_in_episode
    self._episode_ended = ended
    if not ended:
      self._step_count += 1
    self._time_index += 1
    self._current_time_step = self._time_step()
    if self._metrics_writer is not None:
      self._metrics_writer.log_step(a, self._current_time_step)
    return self._current_time_step

  def _upload_convection(self):
    """building.apply_convection() of this step (simulator_flexible_floor_plan.py:156),
    drawn on the host with the reference's generator calls (sbsim_b200/convection.py)."""
    if self._conv is None:
      return
    b = self.building
    shape = (b.height, b.width)
    perm = np.empty((b.n_envs, b.height * b.width), dtype=np.int32)
    for e in range(b.n_envs):
      rooms = self._conv_rooms[0 if len(b.plans) == 1 else e]
      perm[e] = self._conv.gather_index(rooms, shape, self._conv_streams[e])
    self._handle.upload("convection_perm", perm)

  # ---- device-resident I/O (torch tensors as the device-memory container) ----

  def reset_device(self, obs, reward, step_type, discount, stream: int = 0):
    """Like reset() but writes into caller-owned CUDA tensors (asynchronous)."""
    self._before_reset()
    self._handle.reset_device(obs.data_ptr(), reward.data_ptr(), step_type.data_ptr(),
                              discount.data_ptr(), stream)
    self._episode_ended = False
    self._step_count = 0
    self._time_index = 0
    self._current_time_step = True

  def step_device(self, action, obs, reward, step_type, discount, stream: int = 0):
    """Like step() with CUDA tensors in and out; no host round trip.  The caller
    guarantees actions within [-1, 1] (they are not validated on the host)."""
    if self._episode_ended:
      raise RuntimeError("episode has ended; call reset_device()")
    self._upload_convection()
    self._handle.step_device(action.data_ptr(), obs.data_ptr(), reward.data_ptr(),
                             step_type.data_ptr(), discount.data_ptr(), stream)
    ended = self._step_count >= self._num_timesteps_in_episode
    self._episode_ended = ended
    if not ended:
      self._step_count += 1
    self._time_index += 1

  def close(self):
    self._handle.close()

  def render(self, mode: str = "rgb_array"):
    raise NotImplementedError("Rendering not supported yet.")  # environment.py:1362-1363
